"""Device synapse generator, map padding and observation packing (``csrc/env_kernels.cu``) against their CPU
restatements: ``oracle/synapse_oracle.py`` (bit-exact: integer outputs from the same Philox words) and numpy."""
import ctypes

import numpy as np
import pytest

from gym_sted_b200 import _lib
from oracle import synapse_oracle as so


def test_oracle_philox_matches_the_library_block():
    """The oracle's numpy Philox4x32-10 against the library's host-side block function (and through it the Random123
    known answers ``tests/test_abi.py`` checks)."""
    lib = _lib.load()
    rng = np.random.default_rng(0)
    ctrs = rng.integers(0, 2 ** 32, (64, 4), dtype=np.uint64)
    ctrs[0] = 0
    ctrs[1] = 0xFFFFFFFF
    for key in ((0, 0), (0xFFFFFFFF, 0xFFFFFFFF), (0xA4093822, 0x299F31D0)):
        want = so.philox4x32_10(ctrs, key)
        for c, w in zip(ctrs, want):
            out = (ctypes.c_uint32 * 4)()
            lib.sted_philox4x32_10((ctypes.c_uint32 * 4)(*[int(v) for v in c]), (ctypes.c_uint32 * 2)(*key), out)
            assert list(out) == [int(v) for v in w]


def test_oracle_structures_follow_the_generator_configuration():
    """gym_sted/envs/mosted_env.py:80-83: 5 molecules per structure pixel, 3-15 nanodomains of 100-175 molecules at
    least 100 nm = 5 px apart; SURVEY 8d: about 14 % of the frame occupied."""
    frames, coords = so.generate_batch(6, 128, 128, seed=1234)
    dens = []
    for f, c in zip(frames, coords):
        assert set(np.unique(f)) <= {0, 5} | set(range(100, 176))
        assert 4 * 3 <= len(c) <= 4 * 15
        assert all(100 <= f[y, x] <= 175 for y, x in c)
        assert (f > 5).sum() == len(c)
        for cell_y in range(2):
            for cell_x in range(2):
                cc = c[(c[:, 0] // 64 == cell_y) & (c[:, 1] // 64 == cell_x)]
                d = np.linalg.norm(cc[:, None, :] - cc[None, :, :], axis=-1) + 1e9 * np.eye(len(cc))
                assert d.min() >= 5.0
        dens.append((f > 0).mean())
    assert 0.05 < np.mean(dens) < 0.25
    again, _ = so.generate_batch(6, 128, 128, seed=1234)
    assert np.array_equal(frames, again)
    other, _ = so.generate_batch(6, 128, 128, seed=1234, offset=1)
    assert not np.array_equal(frames, other)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,seed,offset,env_base", [(5, 64, 64, 7, 0, 0), (3, 96, 96, 2 ** 40 + 5, 3, 11), (2, 256, 256, 1234, 1, 0),
                                                         (2, 100, 70, 99, 0, 4)])
def test_device_generator_equals_the_oracle_bit_for_bit(B, H, W, seed, offset, env_base):
    import torch
    from gym_sted_b200 import synapse
    gen = synapse.DeviceSynapseGenerator(B, H, W, "cuda", seed=seed, env_base=env_base)
    frames, coords, counts = gen.generate(offset=offset)
    frames, coords, counts = frames.cpu().numpy(), coords.cpu().numpy(), counts.cpu().numpy()
    want_f, want_c = so.generate_batch(B, H, W, seed, offset, env_base)
    assert np.array_equal(frames, want_f)
    for b in range(B):
        assert counts[b] == len(want_c[b])
        assert np.array_equal(coords[b, :counts[b]], want_c[b])
        assert not coords[b, counts[b]:].any()


@pytest.mark.gpu
def test_structures_do_not_depend_on_the_batch_split():
    from gym_sted_b200 import synapse
    whole = synapse.DeviceSynapseGenerator(8, 64, 64, "cuda", seed=5).generate(offset=2)
    part = synapse.DeviceSynapseGenerator(3, 64, 64, "cuda", seed=5, env_base=4).generate(offset=2)
    for w, p in zip(whole, part):
        assert np.array_equal(w[4:7].cpu().numpy(), p.cpu().numpy())


@pytest.mark.gpu
def test_generator_argument_errors_are_reported():
    from gym_sted_b200 import synapse
    with pytest.raises(_lib.StedError):
        synapse.DeviceSynapseGenerator(2, 64, 64, "cuda", n_nanodomains=(3, 40)).generate()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["int32", "float32"])
def test_pad_maps_is_set_roi_max(dtype):
    """Datamap.set_roi(i_ex, "max"): the map sits in a zero frame of K // 2 pixels (gym_sted/utils.py:179-180)."""
    import torch
    lib = _lib.load()
    B, H, W, K = 3, 50, 37, 59
    rng = np.random.default_rng(1)
    roi = rng.integers(0, 200, (B, H, W)).astype(dtype)
    Hp, Wpitch = H + K - 1, (W + K - 1 + 3) & ~3
    out = torch.full((B, Hp, Wpitch), -1.0, dtype=torch.float32, device="cuda")
    d = torch.from_numpy(roi).cuda()
    _lib.check(lib.sted_pad_maps(d.data_ptr(), int(dtype == "int32"), B, H, W, K, out.data_ptr(), None))
    want = np.zeros((B, Hp, Wpitch), dtype=np.float32)
    want[:, K // 2:K // 2 + H, K // 2:K // 2 + W] = roi
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.gpu
def test_pack_observation_casts_like_numpy():
    """contextual_sted_env.py:84-93: numpy.stack((state, conf1, sted), -1).astype(uint16) wraps modulo 2^16."""
    import torch
    lib = _lib.load()
    B, H, W = 2, 9, 13
    rng = np.random.default_rng(2)
    chans = [rng.integers(0, 200000, (B, H, W)).astype(np.float32) for _ in range(3)]
    out = torch.empty((B, H, W, 3), dtype=torch.uint16, device="cuda")
    d = [torch.from_numpy(c).cuda() for c in chans]
    _lib.check(lib.sted_pack_observation(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), B, H, W, out.data_ptr(), None))
    want = np.stack(chans, axis=-1).astype(np.int64).astype(np.uint16)
    assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(96, 96), (64, 80)])
def test_structure_maps_kernel_equals_the_numpy_rendering(shape):
    """``sted_structure_maps`` (one thread per output pixel, rotation / flips undone analytically) against
    ``SyntheticDatamapGenerator.render_numpy`` on the same drawn parameters.  Both evaluate the model in float64; libm
    and CUDA ``exp`` / ``sin`` may differ in the last place, which can move ``floor`` by one molecule at a pixel sitting
    exactly on a step - allowed for at most 1e-4 of the pixels, never by more than one."""
    import torch
    from gym_sted_b200 import envs
    gen = envs.SyntheticDatamapGenerator(shape, seed=11)
    groups = [gen.GROUPS[i % len(gen.GROUPS)] for i in range(64)]
    objects, env = gen.draw_batch(groups)
    want = gen.render_numpy(objects, env)
    gen2 = envs.SyntheticDatamapGenerator(shape, seed=11)
    got = gen2.batch(groups, "cuda:0").cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    diff = np.abs(got - want)
    assert diff.max() <= 1 and (diff > 0).mean() <= 1e-4, (diff.max(), (diff > 0).mean())
    if shape[0] == shape[1]:
        assert (env[:, 2] > 0).any() and (env[:, 3] > 0).any() and (env[:, 4] > 0).any()
    torch.cuda.synchronize()
