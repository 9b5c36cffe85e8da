"""Physical constants, action bounds and objective scales of the reference
(values from ``gym_sted/defaults.py:15-42,102-106,121-133,149-162``; the microscope defaults the
reference leaves to pySTED — objective f/n/NA, STED pulse, detector efficiencies — are SURVEY.md
Appendix A's recalled pySTED defaults)."""
P_STED = 150.0e-3
P_EX = 10.0e-6
PDT = 10.0e-6

LASER_EX = {"lambda_": 635e-9}
LASER_STED = {"lambda_": 750e-9, "zero_residual": 0.01, "anti_stoke": False}
DETECTOR = {"noise": True, "background": 0.5 / (3 * PDT)}
OBJECTIVE = {
    "transmission": {488: 0.84, 535: 0.85, 550: 0.86, 585: 0.85, 575: 0.85, 635: 0.84, 690: 0.82, 750: 0.77,
                     775: 0.75}
}
FLUO = {
    "lambda_": 6.9e-7,
    "qy": 0.65,
    "sigma_abs": {635: 3.2e-21, 750: 3.5e-25},
    "sigma_ste": {750: 3.0e-22},
    "tau": 3.5e-9,
    "tau_vib": 1e-12,
    "tau_tri": 0.0000012,
    "k0": 0,
    "k1": 2.9e-16,
    "b": 1.66,
    "triplet_dynamics_frac": 0,
}

PIXELSIZE = 20e-9

action_spaces = {
    "p_sted": {"low": 0., "high": 350.0e-3},
    "p_ex": {"low": 0., "high": 20.0e-6},
    "pdt": {"low": 1.0e-6, "high": 100.0e-6},
}

bounds_dict = {
    "SNR": {"low": 0.20, "high": float("inf")},
    "Bleach": {"low": -float("inf"), "high": 0.5},
    "Resolution": {"low": 0, "high": 100},
    "NbNanodomains": {"low": 0, "high": float("inf")},
}
scales_dict = {
    "SNR": {"low": 0, "high": 1},
    "Bleach": {"low": 0, "high": 1},
    "Resolution": {"low": 40, "high": 180},
    "NbNanodomains": {"low": 0, "high": 1},
}

fluorescence_criterions = {
    "bleach": {"p_ex": [1.0e-6, 10.0e-6], "p_sted": [100e-3, 200e-3], "pdt": [5.0e-6, 75.0e-6], "target": [0.1, 0.9]},
    "signal": {"p_ex": [1.0e-6, 10.0e-6], "p_sted": 0., "pdt": 10.0e-6, "target": [10., 400.]},
}

# Frozen fluorophore parameter sets: the (b, k1, sigma_abs[635]) triples fitted by the reference's
# BleachSampler("uniform") for its four routines (values of gym_sted/routines/routines.json:2-73;
# everything else in a routine equals FLUO above).
ROUTINE_FITS = {
    "high-signal_high-bleach": {"b": 1.738794634421001, "k1": 3.073224847565423e-16, "sigma_abs_635": 3.2653303994069316e-21},
    "high-signal_low-bleach": {"b": 1.6184803383889534, "k1": 3.626623963067252e-16, "sigma_abs_635": 3.2482521692221336e-21},
    "low-signal_high-bleach": {"b": 1.7933283778872158, "k1": 3.1902952522748453e-16, "sigma_abs_635": 4.7522773364092475e-22},
    "low-signal_low-bleach": {"b": 1.6770989185855834, "k1": 3.323874455981149e-16, "sigma_abs_635": 4.403015011667726e-22},
}


def load_routine(routine_id: str) -> dict:
    """Fluorophore dict of a named routine (loader semantics of ``gym_sted/utils.py:302-314``)."""
    fit = ROUTINE_FITS[routine_id]
    routine = {k: (dict(v) if isinstance(v, dict) else v) for k, v in FLUO.items()}
    routine["b"], routine["k1"] = fit["b"], fit["k1"]
    routine["sigma_abs"][635] = fit["sigma_abs_635"]
    return routine
