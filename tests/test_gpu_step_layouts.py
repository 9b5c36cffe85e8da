"""The two schedules of ``VectorContextualMOSTED.step`` (``envs.py``: "narrow" = acquisitions and decorrelation on the
caller's stream, "wide" = acquisitions on a high-priority stream with three side streams) launch the same kernels on
the same inputs: observations, objectives, F1 and the next structures must be identical, step after step."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gym_sted_b200 import envs  # noqa: E402

pytestmark = pytest.mark.gpu


def _episode(layout, monkeypatch, env_id, steps=3, B=6, size=64):
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")
    monkeypatch.setenv("STED_B200_STEP_LAYOUT", layout)
    vec = envs.make_vec(env_id, B, compute_f1=True, seed=11, image_size=size)
    assert vec.wide_step == (layout == "wide")
    (img, v), _ = vec.reset(seed=5)
    out = [img.copy(), v.copy()]
    rng = np.random.default_rng(3)
    for _ in range(steps):
        a = rng.uniform(vec.action_space.low, vec.action_space.high, (B, 3)).astype(np.float32)
        (img, v), reward, done, _, info = vec.step(a)
        out += [img.copy(), v.copy(), reward.copy(), info["mo_objs"].copy(), info["f1-score"].copy(),
                info["bleached"].cpu().numpy(), info["sted_image"].cpu().numpy(), info["conf2"].cpu().numpy(),
                np.concatenate([c.reshape(-1) for c in info["nanodomains-coords"]])]
    return out


@pytest.mark.parametrize("env_id", ["ContextualMOSTED-easy-hslb-v0", "SequenceMOSTED-easy-v0"])
def test_wide_and_narrow_steps_are_identical(monkeypatch, env_id):
    if env_id not in envs.REGISTRY:
        pytest.skip(f"{env_id} not registered")
    narrow = _episode("narrow", monkeypatch, env_id)
    wide = _episode("wide", monkeypatch, env_id)
    assert len(narrow) == len(wide)
    for i, (x, y) in enumerate(zip(narrow, wide)):
        np.testing.assert_array_equal(x, y, err_msg=f"item {i}")


def test_layout_is_chosen_by_step_size(monkeypatch):
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")
    monkeypatch.delenv("STED_B200_STEP_LAYOUT", raising=False)
    assert not envs.make_vec("ContextualMOSTED-easy-hslb-v0", 4, seed=0, image_size=64).wide_step
    assert envs.make_vec("ContextualMOSTED-easy-hslb-v0", 64, seed=0, image_size=256).wide_step
    monkeypatch.setenv("STED_B200_STEP_LAYOUT", "sideways")
    with pytest.raises(ValueError):
        envs.make_vec("ContextualMOSTED-easy-hslb-v0", 4, seed=0, image_size=64)


def test_fluorophore_fits_queued_ahead_give_the_same_episodes(monkeypatch):
    """``ContextualMOSTED-hard-v0``: the fits of the next episode are queued at a reset and run under the episode's
    steps (``STED_B200_PREFETCH_FLUO``); criteria, fitted parameters and rewards are those of fitting at reset time."""
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (no CPU fallback exists)")
    runs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("STED_B200_PREFETCH_FLUO", flag)
        vec = envs.make_vec("ContextualMOSTED-hard-v0", 5, compute_f1=True, seed=21, image_size=64)
        assert vec.prefetch_fluorophores == (flag == "1")
        rng = np.random.default_rng(8)
        out = []
        for episode in range(3):
            vec.reset(seed=4 if episode == 0 else None)           # unseeded resets consume the queued fits
            out.append(np.array([c["signal"]["target"] for c in vec.bleach_sampler.criteria]))
            out.append(vec.bm.d_fluo.cpu().numpy().copy())           # the fitted StedFluoParams of every environment
            for _ in range(2):
                a = rng.uniform(vec.action_space.low, vec.action_space.high, (5, 3)).astype(np.float32)
                (img, v), reward, _, _, info = vec.step(a)
                out += [img.copy(), reward.copy(), info["mo_objs"].copy()]
        runs.append(out)
    for i, (x, y) in enumerate(zip(*runs)):
        np.testing.assert_array_equal(x, y, err_msg=f"item {i}")
