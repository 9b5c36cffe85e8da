"""Drop-in objective classes: the reference's ``gym_sted.rewards.objectives`` interface on the CUDA reward kernels.

``gym_sted/envs/mosted_env.py:86-97`` builds ``{"Resolution": objectives.Resolution(pixelsize, res_cap), "Bleach":
objectives.Bleach(), "SNR": objectives.Signal_Ratio(75), "NbNanodomains": objectives.NumberNanodomains()}`` and hands it
to ``rewards.MORewardCalculator`` / ``rewards.NanodomainsRewardCalculator``; every ``env.step`` then calls
``evaluate(sted_stack, confocal_init, confocal_end, sted_fg, confocal_fg)`` on each (``contextual_sted_env.py:58-59``).
The classes here have the same names, constructor arguments, ``label`` / ``select_optimal`` attributes, ``evaluate``
signatures, return values (``None`` for a negative SNR, ``0`` without STED foreground, the capped resolution in nm, the
F1 score) and take the same numpy images and masks - the arithmetic runs in ``objectives_kernel`` (masks supplied by
the caller: ``sted_objectives_masked``), ``decorr_kernel`` and ``nanodomain_kernel`` / ``gaussfit_kernel`` /
``f1_match_kernel``.  A maintainer swaps ``from gym_sted.rewards import objectives`` for
``from gym_sted_b200 import objectives``; the vectorised environments of ``envs.py`` do not go through
these one-image calls.  No CPU fallback: without a CUDA device every ``evaluate`` raises ``StedError``.
"""
from abc import ABC, abstractmethod

import numpy as np

from . import _lib, defaults


def _torch():
    import torch
    if _lib.device_count() == 0:
        raise _lib.StedError(-2, "no CUDA device: gym_sted_b200 has no CPU fallback")
    return torch


def _image(x):
    """One (H, W) image (array, list holding one, or tensor) -> (1, H, W) float32 CUDA tensor."""
    torch = _torch()
    if torch.is_tensor(x):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    t = t.to("cuda", torch.float32)
    return t.reshape(1, *t.shape[-2:])


def _mask(x):
    torch = _torch()
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np.uint8)))
    t = t.to("cuda").to(torch.uint8)
    return t.reshape(1, *t.shape[-2:]).contiguous()


def masked_objectives(conf1, sted, conf2, fg_s, fg_c):
    """(Bleach, SNR, flags) of one acquisition for caller-supplied masks (``sted_objectives_masked``); SNR is NaN where
    numpy's percentile of an empty selection is."""
    torch = _torch()
    c1, st, c2 = _image(conf1), _image(sted), _image(conf2)
    mc, ms = _mask(fg_c), _mask(fg_s)
    objs = torch.empty((1, 2), dtype=torch.float64, device=c1.device)
    flags = torch.empty((1,), dtype=torch.int32, device=c1.device)
    scratch = torch.empty((1, 2, 65536), dtype=torch.int32, device=c1.device)
    lib = _lib.load()
    _lib.check(lib.sted_objectives_masked(c1.data_ptr(), st.data_ptr(), c2.data_ptr(), 1, c1.shape[1], c1.shape[2],
                                          mc.data_ptr(), ms.data_ptr(), objs.data_ptr(), flags.data_ptr(),
                                          scratch.data_ptr(), torch.cuda.current_stream(c1.device).cuda_stream))
    o, f = objs.cpu().numpy()[0], int(flags.cpu().numpy()[0])
    if f & 2:
        raise _lib.StedError(-4, "objectives_kernel: image value range beyond 65536 histogram bins")
    return float(o[0]), float(o[1]), f


def get_foreground(img):
    """``gym_sted.utils.get_foreground`` (``utils.py:16-24``): Otsu mask of one image, computed by ``objectives_kernel``
    (bit-identical to skimage's threshold on integer images); returns a bool array."""
    from . import rewards
    t = _image(img)
    _, fg_c, _, _, _ = rewards.foreground_objectives(t, t, t)
    return fg_c[0].cpu().numpy()


class Objective(ABC):
    """``objectives.py:28-55``: ``label``, ``select_optimal`` and ``evaluate``."""

    label = ""
    select_optimal = None

    @abstractmethod
    def evaluate(self, sted_stack, confocal_init, confocal_end, sted_fg, confocal_fg):
        raise NotImplementedError

    def mirror_ticks(self, ticks):
        return ticks


class Signal_Ratio(Objective):
    """``objectives.py:57-100``: (P_q(sted[fg_s]) - mean(sted[~fg_s])) / P_q(conf1[fg_c]); ``0`` without STED foreground,
    ``None`` when negative.  The kernel is compiled for the percentile every gym-sted environment uses (75)."""

    def __init__(self, percentile):
        if float(percentile) != 75.0:
            raise ValueError("objectives_kernel computes the 75th percentile (the only one gym-sted instantiates)")
        self.label = "Signal Ratio"
        self.select_optimal = np.argmax
        self.percentile = percentile

    def evaluate(self, sted_stack, confocal_init, confocal_end, sted_fg, confocal_fg):
        if not np.any(np.asarray(sted_fg.cpu() if hasattr(sted_fg, "cpu") else sted_fg)):
            return 0
        _, snr, flags = masked_objectives(confocal_init, sted_stack[0], confocal_end, sted_fg, confocal_fg)
        return None if flags & 1 else snr


class Bleach(Objective):
    """``objectives.py:102-111``: (mean(conf1[fg_c]) - mean(conf2[fg_c])) / mean(conf1[fg_c])."""

    def __init__(self):
        self.label = "Bleach"
        self.select_optimal = np.argmin

    def evaluate(self, sted_stack, confocal_init, confocal_end, sted_fg, confocal_fg):
        bleach, _, _ = masked_objectives(confocal_init, sted_stack[0], confocal_end, sted_fg, confocal_fg)
        return bleach


class Resolution(Objective):
    """``objectives.py:113-291``: image-decorrelation resolution of the STED image in nm, capped at ``res_cap``."""

    def __init__(self, pixelsize, res_cap=250):
        self.label = "Resolution (nm)"
        self.select_optimal = np.argmin
        self.pixelsize = pixelsize
        self.res_cap = res_cap

    def evaluate(self, sted_stack, confocal_init, confocal_end, sted_fg, confocal_fg):
        from . import rewards
        return float(rewards.resolution(_image(sted_stack[0]), pixelsize=self.pixelsize, res_cap=self.res_cap)[0])


class NumberNanodomains(Objective):
    """``objectives.py:387-430``: F1 score of the nanodomains detected in the STED image against ``synapse``'s ground
    truth.  As in the reference the first argument is the IMAGE (``NanodomainsRewardCalculator`` passes it as is)."""

    def __init__(self, threshold=2):
        self.label = "NbNanodomains"
        self.select_optimal = None
        self.threshold = threshold

    def evaluate(self, sted_stack, confocal_init, confocal_end, sted_fg, confocal_fg, *args, **kwargs):
        from . import rewards
        synapse = kwargs.get("synapse")
        pixelsize = synapse.datamap_pixelsize_nm * 1e-9
        truth = np.asarray(synapse.nanodomains_coords, dtype=np.float64).reshape(-1, 2)
        return float(rewards.nanodomain_f1_batch(_image(sted_stack), [truth], pixelsize=pixelsize,
                                                 threshold=self.threshold)[0])


# ---------------------------------------------------------------------------- gym_sted/rewards/rewards.py
class RewardCalculator:
    """``rewards.py:4-38``."""

    def __init__(self, objs, *args, **kwargs):
        self.objectives = objs

    def evaluate(self, sted_image, conf1, conf2, fg_s, fg_c):
        raise NotImplementedError

    def rescale(self, value, obj_name):
        lo, hi = self.scales[obj_name]["low"], self.scales[obj_name]["high"]
        return (value - lo) / (hi - lo)


class MORewardCalculator(RewardCalculator):
    """``rewards.py:40-87``: the list of objective values, in the order of the ``objs`` dict.  When the dict holds this
    module's Bleach and Signal_Ratio both, the two come out of ONE kernel launch."""

    def __init__(self, objs, *args, **kwargs):
        super().__init__(objs)
        self.scales = kwargs.get("scales", None)

    def _values(self, sted_image, conf1, conf2, fg_s, fg_c):
        shared = None
        out = []
        for obj in self.objectives.values():
            if type(obj) in (Bleach, Signal_Ratio):
                if shared is None:
                    shared = masked_objectives(conf1, sted_image, conf2, fg_s, fg_c)
                if type(obj) is Bleach:
                    out.append(shared[0])
                elif not np.any(np.asarray(fg_s.cpu() if hasattr(fg_s, "cpu") else fg_s)):
                    out.append(0)
                else:
                    out.append(None if shared[2] & 1 else shared[1])
            else:
                out.append(obj.evaluate([sted_image], conf1, conf2, fg_s, fg_c))
        return out

    def evaluate(self, sted_image, conf1, conf2, fg_s, fg_c):
        return self._values(sted_image, conf1, conf2, fg_s, fg_c)

    def evaluate_rescale(self, sted_image, conf1, conf2, fg_s, fg_c):
        return [self.rescale(v, name) for v, name in zip(self._values(sted_image, conf1, conf2, fg_s, fg_c), self.objectives)]


class NanodomainsRewardCalculator(RewardCalculator):
    """``rewards.py:200-236``: the rescaled F1 score of the nanodomain detection."""

    def __init__(self, objs, *args, **kwargs):
        super().__init__(objs)
        self.scales = kwargs.get("scales")

    def evaluate(self, sted_stack, confocal_init, confocal_end, sted_fg, confocal_fg, threshold=2, *args, **kwargs):
        return self.rescale(self.objectives["NbNanodomains"].evaluate(sted_stack, confocal_init, confocal_end, sted_fg,
                                                                       confocal_fg, *args, **kwargs), "NbNanodomains")


def default_objectives(pixelsize=defaults.PIXELSIZE, res_cap=None):
    """The objective dict of ``mosted_env.py:86-92`` built from this module's classes."""
    from . import rewards
    return {"Resolution": Resolution(pixelsize, res_cap if res_cap is not None else rewards.RES_CAP),
            "Bleach": Bleach(), "SNR": Signal_Ratio(75)}


def install_into_gym_sted():
    """Zero-source-change form of the swap: ``gym_sted/defaults.py:114-120`` holds ONE dict of objective instances
    (``obj_dict``) that every environment module imports by name, so replacing its entries in place puts the CUDA
    objectives behind every environment constructed afterwards (the reference's own ``MORewardCalculator`` /
    ``NanodomainsRewardCalculator`` then call ``evaluate`` on these objects).  ``Squirrel`` (not on the path of the
    BASELINE environments) is left alone.  Returns the replaced entries so that a caller can put them back."""
    import gym_sted.defaults as ref_defaults
    table = ref_defaults.obj_dict
    old = {k: table[k] for k in ("SNR", "Bleach", "Resolution", "NbNanodomains")}
    table["SNR"] = Signal_Ratio(getattr(old["SNR"], "percentile", 75))
    table["Bleach"] = Bleach()
    table["Resolution"] = Resolution(old["Resolution"].pixelsize, old["Resolution"].res_cap)
    table["NbNanodomains"] = NumberNanodomains(getattr(old["NbNanodomains"], "threshold", 2))
    return old
