"""Trajectory kernels (sg_traj_* in include/softgrip.h): the trainer-side noise augmentation
(ref: functions/optimization.py:6-14) and channel statistics (ref: functions/utils.py:39-40).

CPU tests: the numpy oracle against the published Philox4x32-10 known-answer vectors, and the kernel SOURCE compiled under
the SIMT emulator (tests/simt) against the oracle.  `-m gpu` tests repeat the comparison on the device through the Python
mirror (soft-grip_b200/functions) and add the size-independent properties at BASELINE configs[2] size."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, pkg
from oracle import traj_oracle as to

sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))

# fp32 Box-Muller on the device vs float64 Box-Muller of the same uniforms in the oracle.  The device evaluates it on the
# special-function unit (csrc/sg_traj.cuh box_muller): sin/cos.approx 2^-20.9 = 5.1e-7 absolute on an angle that itself carries
# <= 3e-7 of fp32 rounding, times r <= 6.76, plus the 2^-22 relative error of lg2.approx -> 8e-6 absolute on the standard
# normal in the worst corner (r at its maximum), about 1e-6 typically.  (The emulator build runs the libm formula.)
Z_TOL = 8e-6


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def test_oracle_philox_matches_the_published_known_answer_vectors():
    for ctr, key, want in to.PHILOX_KAT:
        got = to.philox4x32_10(np.array(ctr), key)
        assert [int(v) for v in got] == list(want)
    # vectorised form == element-wise form
    ctr = np.array([[i, 7 - i, 0, 0] for i in range(8)])
    many = to.philox4x32_10(ctr, (0xdeadbeef, 0x12345678))
    for i in range(8):
        assert (many[i] == to.philox4x32_10(ctr[i], (0xdeadbeef, 0x12345678))).all()


def test_oracle_normals_are_standard_and_channelwise_sigmas_follow_the_reference():
    z = to.standard_normals(1 << 20, seed=3)
    assert abs(z.mean()) < 4e-3 and abs(z.std() - 1) < 3e-3
    assert abs((z ** 3).mean()) < 2e-2 and abs((z ** 4).mean() - 3) < 5e-2
    x = np.zeros((64, 200, 12))
    y = to.noised_modality(x, seed=11)
    sd = y.std(axis=(0, 1))
    assert np.allclose(sd[:6], 0.7, rtol=0.03) and np.allclose(sd[6:], 0.06, rtol=0.03)   # ref: optimization.py:8-12
    assert (to.noised_modality(x, seed=11) == y).all() and (to.noised_modality(x, seed=12) != y).any()


@pytest.fixture(scope="module")
def emulib():
    import emu
    return emu.lib()


def _traj(n, t, c, dtype, seed=0, offset=True):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, t, c)) * np.linspace(0.5, 30.0, c)
    if offset:
        x += np.linspace(-9.81, 250.0, c)          # gravity-like offsets: the cancellation case of a one-pass variance
    return np.ascontiguousarray(x.astype(dtype))


@pytest.mark.parametrize("dtype,nchan", [(np.float32, 12), (np.float64, 12), (np.float32, 24), (np.float32, 4)])
def test_emulated_noise_kernel_matches_the_oracle(emulib, dtype, nchan):
    n = 37 if nchan == 12 else 5                   # 22 200 quads > the emulated grid of 4 096 threads: grid-stride path
    x = _traj(n, 200, nchan, dtype, seed=1)
    out = np.empty_like(x)
    prec = 32 if dtype == np.float32 else 64
    seed = 0x1234567890ABCDEF
    assert emulib.sg_traj_add_noise(_vp(x), _vp(out), x.size // nchan, 0, nchan, nchan // 2, 0.7, 0.06, seed, None, None, prec, 0, None) == 0
    want = to.noised_modality(x, seed)
    sig = np.where(np.arange(nchan) < nchan // 2, 0.7, 0.06)
    tol = sig * Z_TOL + (np.abs(want) * (2e-7 if prec == 32 else 1e-15))
    assert (np.abs(out - want) <= tol).all()
    # in place (the reference's `acc += ...`) gives the same bits
    y = x.copy()
    assert emulib.sg_traj_add_noise(_vp(y), _vp(y), x.size // nchan, 0, nchan, nchan // 2, 0.7, 0.06, seed, None, None, prec, 0, None) == 0
    assert (y == out).all()
    # a shard noised on its own with its row offset equals the same rows of the whole tensor (per-launch / per-GPU shards)
    rows, cut = x.size // nchan, 3 * 200 + 7
    tail = np.ascontiguousarray(x.reshape(rows, nchan)[cut:])
    ytail = np.empty_like(tail)
    assert emulib.sg_traj_add_noise(_vp(tail), _vp(ytail), rows - cut, cut, nchan, nchan // 2, 0.7, 0.06, seed, None, None, prec, 0, None) == 0
    assert (ytail == out.reshape(rows, nchan)[cut:]).all()
    assert (np.abs(ytail - to.noised_modality(tail, seed, first_row=cut)) <= tol.reshape(rows, nchan)[cut:]).all()
    # sigma 0 is the identity; another seed is another draw
    assert emulib.sg_traj_add_noise(_vp(x), _vp(y), x.size // nchan, 0, nchan, nchan // 2, 0.0, 0.0, seed, None, None, prec, 0, None) == 0
    assert (y == x).all()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(y), x.size // nchan, 0, nchan, nchan // 2, 0.7, 0.06, seed + 1, None, None, prec, 0, None) == 0
    assert (y != out).mean() > 0.99


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_emulated_noise_kernel_fused_standardisation(emulib, dtype):
    x = _traj(9, 200, 12, dtype, seed=2)
    mean, std = to.channel_mean_std(x)
    mean, std = np.ascontiguousarray(mean.reshape(-1)), np.ascontiguousarray(std.reshape(-1))
    out = np.empty_like(x)
    prec = 32 if dtype == np.float32 else 64
    assert emulib.sg_traj_add_noise(_vp(x), _vp(out), x.size // 12, 0, 12, 6, 0.7, 0.06, 99, _vp(mean), _vp(std), prec, 0, None) == 0
    want = to.noised_modality(x, 99, mean=mean, std=std)                # ref: optimization.py:33,38
    tol = (np.where(np.arange(12) < 6, 0.7, 0.06) * Z_TOL) / std + (np.abs(want) + np.abs(mean / std)) * (4e-7 if prec == 32 else 1e-14)
    assert (np.abs(out - want) <= tol).all()


@pytest.mark.parametrize("dtype,shape", [(np.float32, (37, 200, 12)), (np.float64, (37, 200, 12)), (np.float32, (3, 50, 24)),
                                         (np.float64, (1, 1, 12)), (np.float32, (2, 3, 52)), (np.float32, (5, 7, 4))])
def test_emulated_channel_stats_match_numpy(emulib, dtype, shape):
    x = _traj(*shape, dtype, seed=4)
    nchan = shape[-1]
    nrows = x.size // nchan
    prec = 32 if dtype == np.float32 else 64
    nbytes = emulib.sg_traj_stats_workspace_bytes(nrows, nchan, 0)
    assert nbytes > 0
    ws = np.full(nbytes // 8, np.nan)
    mean, std = np.empty(nchan), np.empty(nchan)
    assert emulib.sg_traj_channel_stats(_vp(x), nrows, nchan, prec, 0, _vp(mean), _vp(std), _vp(ws), nbytes, None) == 0
    wm, wsd = to.channel_mean_std(x)                                     # ref: functions/utils.py:39-40
    assert np.allclose(mean, wm.reshape(-1), rtol=1e-12, atol=1e-12)
    assert np.allclose(std, wsd.reshape(-1), rtol=1e-10, atol=1e-12)
    mean2, std2 = np.empty(nchan), np.empty(nchan)
    assert emulib.sg_traj_channel_stats(_vp(x), nrows, nchan, prec, 0, _vp(mean2), _vp(std2), _vp(ws), nbytes, None) == 0
    assert (mean2 == mean).all() and (std2 == std).all()                 # fixed summation order


def _touch(W, T, seed, p_finger=0.3):
    """Random per-row contact words: finger bits 0/1, TOUCH_ANY (bit 30) when there is any contact at all."""
    rng = np.random.default_rng(seed)
    f = (rng.random((W, T)) < p_finger).astype(np.int32) | ((rng.random((W, T)) < p_finger).astype(np.int32) << 1)
    anyc = (f != 0) | (rng.random((W, T)) < 0.3)
    t = f | (anyc.astype(np.int32) << 30)
    t[0] = 0                       # a world that never touches
    t[1] = 3 | (1 << 30)           # a world that always touches with both fingers
    return np.ascontiguousarray(t)


@pytest.mark.parametrize("dtype,T,nchan", [(np.float32, 200, 12), (np.float64, 200, 12), (np.float32, 33, 24), (np.float32, 1, 4)])
@pytest.mark.parametrize("mode", ["intended", "reference"])
def test_emulated_mask_contact_matches_the_literal_restatement(emulib, dtype, T, nchan, mode):
    W = 21                                                    # 16 warps in the emulated grid: grid-stride over worlds
    x = _traj(W, T, nchan, dtype, seed=6) + 1000.0            # no accidental zeros
    touch = _touch(W, T, seed=7, p_finger=0.08 if mode == "reference" else 0.6)
    prec = 32 if dtype == np.float32 else 64
    y = x.copy()
    left = np.full(W, 3, dtype=np.int32)
    left[5] = 0                                               # a world whose list ran empty in an earlier episode
    lists = [[k for k in range(2) if (int(v) >> k) & 1] for v in left]
    m = 1 if mode == "reference" else 0
    assert emulib.sg_traj_mask_contact(_vp(y), _vp(touch), W, T, nchan, 3, 1 << 30, m, _vp(left) if m else None, prec, 0, None) == 0
    want = to.mask_contact(x, touch, 2, 1 << 30, mode, fingers_left=lists)
    assert (y == want).all()
    assert 0 < (want == 0).all(axis=-1).mean() < 1            # the case masks some rows and keeps some
    if m:
        assert [[k for k in range(2) if (int(v) >> k) & 1] for v in left] == lists      # the carried list state
        # second episode continues with the carried lists
        y2 = x.copy()
        assert emulib.sg_traj_mask_contact(_vp(y2), _vp(touch), W, T, nchan, 3, 1 << 30, 1, _vp(left), prec, 0, None) == 0
        assert (y2 == to.mask_contact(x, touch, 2, 1 << 30, mode, fingers_left=lists)).all()
        # without carried state every call starts from the full list
        y3 = x.copy()
        assert emulib.sg_traj_mask_contact(_vp(y3), _vp(touch), W, T, nchan, 3, 1 << 30, 1, None, prec, 0, None) == 0
        assert (y3 == to.mask_contact(x, touch, 2, 1 << 30, mode)).all()


def test_batched_contact_flag_agrees_with_the_row_scan():
    """BatchedManEnv._contact_flag (the per-step torch expression) and the rollout mask use the same semantics."""
    touch = _touch(6, 40, seed=9, p_finger=0.1)
    x = np.ones((6, 40, 4))
    want = to.mask_contact(x, touch, 2, 1 << 30, "reference")
    left = np.full(6, 3)
    for t in range(40):
        was_empty = left == 0
        left = left & ~(touch[:, t] & 3)
        flag = np.where(was_empty, (touch[:, t] & (1 << 30)) != 0, left == 0)     # batched.py _contact_flag
        assert (flag == (want[:, t, 0] != 0)).all()


def test_trajectory_kernel_source_is_clean_under_asan_and_ubsan(tmp_path):
    """tests/simt/asan_traj.cpp: the three kernels over ragged shapes / grids on exact-size heap buffers with
    -fsanitize=address,undefined -- no out-of-bounds access, no undefined behaviour in the kernel source."""
    import subprocess
    exe = str(tmp_path / "asan_traj")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-I" + os.path.join(ROOT, "tests", "simt"), "-I" + os.path.join(ROOT, "soft-grip_b200", "csrc"), "-Wno-unknown-pragmas",
           "-o", exe, os.path.join(ROOT, "tests", "simt", "asan_traj.cpp")]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and "sanitize" in built.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert built.returncode == 0, built.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_stack_use_after_return=0:detect_leaks=0")
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "asan driver done" in out.stdout, (out.stdout[-500:], out.stderr[-3000:])
    assert "ERROR: AddressSanitizer" not in out.stderr and "runtime error" not in out.stderr, out.stderr[-3000:]


def test_trajectory_entry_points_reject_bad_arguments(emulib):
    x = np.zeros((4, 12), dtype=np.float32)
    bad = np.zeros(4 * 12 + 1, dtype=np.float32)[1:]                     # 4-byte aligned only
    m = np.zeros(12)
    err = lambda: emulib.sg_last_error().decode()
    assert emulib.sg_traj_add_noise(None, _vp(x), 4, 0, 12, 6, 0.7, 0.06, 0, None, None, 32, 0, None) < 0 and "null" in err()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(x), 4, 0, 10, 5, 0.7, 0.06, 0, None, None, 32, 0, None) < 0 and "multiple of 4" in err()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(x), 4, 0, 12, 13, 0.7, 0.06, 0, None, None, 32, 0, None) < 0 and "nacc" in err()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(x), 4, -1, 12, 6, 0.7, 0.06, 0, None, None, 32, 0, None) < 0 and "first_row" in err()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(x), 4, 0, 12, 6, -1.0, 0.06, 0, None, None, 32, 0, None) < 0 and "sigma" in err()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(x), 4, 0, 12, 6, 0.7, 0.06, 0, _vp(m), None, 32, 0, None) < 0 and "together" in err()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(x), 4, 0, 12, 6, 0.7, 0.06, 0, None, None, 16, 0, None) < 0 and "precision" in err()
    assert emulib.sg_traj_add_noise(_vp(bad), _vp(x), 4, 0, 12, 6, 0.7, 0.06, 0, None, None, 32, 0, None) < 0 and "aligned" in err()
    assert emulib.sg_traj_add_noise(_vp(x), _vp(x), 0, 0, 12, 6, 0.7, 0.06, 0, None, None, 32, 0, None) == 0           # empty is a no-op
    ws = np.zeros(8)
    assert emulib.sg_traj_channel_stats(_vp(x), 4, 12, 32, 0, _vp(m), _vp(m), _vp(ws), 8, None) < 0 and "workspace" in err()
    assert emulib.sg_traj_channel_stats(_vp(x), 0, 12, 32, 0, _vp(m), _vp(m), _vp(ws), 64, None) < 0 and "one row" in err()
    assert emulib.sg_traj_stats_workspace_bytes(4, 7, 0) < 0
    t = np.zeros(4, dtype=np.int32)
    assert emulib.sg_traj_mask_contact(_vp(x), None, 4, 1, 12, 3, 1 << 30, 0, None, 32, 0, None) < 0 and "touch" in err()
    assert emulib.sg_traj_mask_contact(_vp(x), _vp(t), 4, 1, 12, 3, 1 << 30, 2, None, 32, 0, None) < 0 and "mode" in err()
    assert emulib.sg_traj_mask_contact(_vp(x), _vp(t), 4, 1, 12, 3, 1, 0, None, 32, 0, None) < 0 and "any_bit" in err()
    assert emulib.sg_traj_mask_contact(_vp(x), _vp(t), 0, 1, 12, 3, 1 << 30, 0, None, 32, 0, None) == 0


def test_python_mirror_fails_loudly_without_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    fn = pkg("functions")
    lib = pkg("_lib")
    with pytest.raises(lib.SoftGripError):
        fn.noised_modality(torch.zeros(2, 200, 12))
    with pytest.raises(lib.SoftGripError):
        fn.channel_mean_std(torch.zeros(2, 200, 12))
    # the C entry points themselves refuse to run without a device
    L = lib.lib()
    x = np.zeros((4, 12), dtype=np.float32)
    assert L.sg_traj_add_noise(_vp(x), _vp(x), 4, 0, 12, 6, 0.7, 0.06, 0, None, None, 32, 0, None) < 0
    assert b"no CUDA device" in L.sg_last_error()


# ---- on the device -----------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_gpu_noise_and_stats_match_the_oracle(torch_cuda, dtype):
    torch = torch_cuda
    fn = pkg("functions")
    npdt = np.float32 if dtype == "float32" else np.float64
    x = _traj(301, 200, 12, npdt, seed=5)                                # 180 600 quads: ragged last CTA
    xd = torch.from_numpy(x).cuda()
    seed = 2 ** 63 + 12345
    out = fn.noised_modality(xd, seed=seed)
    want = to.noised_modality(x, seed)
    sig = np.where(np.arange(12) < 6, 0.7, 0.06)
    tol = sig * Z_TOL + np.abs(want) * (2e-7 if dtype == "float32" else 1e-15)
    assert (np.abs(out.cpu().numpy() - want) <= tol).all()
    assert (xd.cpu().numpy() == x).all()                                 # input untouched unless out is data
    inplace = fn.noised_modality(xd.clone(), seed=seed)
    y = xd.clone()
    fn.noised_modality(y, seed=seed, out=y)
    assert torch.equal(y, out) and torch.equal(inplace, out)
    mean, std = fn.channel_mean_std(xd)
    wm, wsd = to.channel_mean_std(x)
    assert mean.shape == (1, 1, 12) and std.shape == (1, 1, 12)
    assert np.allclose(mean.cpu().numpy(), wm, rtol=1e-12, atol=1e-12) and np.allclose(std.cpu().numpy(), wsd, rtol=1e-10, atol=1e-12)
    fused = fn.noised_modality(xd, seed=seed, mean=mean, std=std)
    wantf = to.noised_modality(x, seed, mean=wm, std=wsd)
    tolf = (sig * Z_TOL) / wsd.reshape(-1) + (np.abs(wantf) + np.abs(wm / wsd).reshape(-1)) * (4e-7 if dtype == "float32" else 1e-14)
    assert (np.abs(fused.cpu().numpy() - wantf) <= tolf).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["intended", "reference"])
def test_gpu_rollout_mask_contact_matches_the_restatement(torch_cuda, mode):
    """A real rollout's touch words: the device mask equals the literal restatement, and (reference mode) the carried
    finger lists continue into a second episode."""
    torch = torch_cuda
    from conftest import blob_path
    batched = pkg("batched")
    env = batched.BatchedManEnv(blob_path("softbox"), 96, dtype=torch.float32, seed=3, contact_mode=mode)
    sched = batched.default_schedule(2, n_settle=4, n_iter=60, open_close_div=30)
    lists = [[0, 1] for _ in range(96)]
    for episode in range(2):
        traj, k, st, touch = env.rollout(schedule=sched, return_touch=True)
        raw = traj.cpu().numpy().copy()
        env.mask_contact(traj, touch)
        want = to.mask_contact(raw, touch.cpu().numpy(), 2, batched.TOUCH_ANY, mode, fingers_left=lists if mode == "reference" else None)
        assert (traj.cpu().numpy() == want).all()
        zero_rows = (want == 0).all(axis=-1)
        assert zero_rows[:, :4].all()                                   # nothing touches during the settle rows
        assert not zero_rows.all()                                      # the squeeze makes contact
    if mode == "reference":
        assert env._fingers_left.cpu().numpy().tolist() == [sum(1 << k for k in l) for l in lists]


@pytest.mark.gpu
def test_gpu_trajectory_kernels_at_full_size_properties(torch_cuda):
    """BASELINE configs[2] size (65 536 worlds x 200 rows x 12 channels, fp32): properties that need no oracle run."""
    torch = torch_cuda
    fn = pkg("functions")
    W = 65536
    base = torch.linspace(-3, 3, 12, device="cuda", dtype=torch.float32)
    x = base.expand(W, 200, 12).contiguous()
    y = fn.noised_modality(x, seed=7)
    d = (y - x).double()
    sd = d.std(dim=(0, 1)).cpu().numpy()
    assert np.allclose(sd[:6], 0.7, rtol=2e-3) and np.allclose(sd[6:], 0.06, rtol=2e-3)
    assert float(d.mean(dim=(0, 1)).abs().max()) < 1.5e-3            # 0.7 / sqrt(1.3e7) = 2e-4 per channel
    # sharding invariance: each half noised on its own (the second with its row offset) equals the halves of the whole
    yh = fn.noised_modality(x[: W // 2].contiguous(), seed=7)
    assert torch.equal(yh, y[: W // 2])
    yt = fn.noised_modality(x[W // 2:].contiguous(), seed=7, first_row=(W // 2) * 200)
    assert torch.equal(yt, y[W // 2:])
    mean, std = fn.channel_mean_std(y)
    m64, s64 = y.double().mean(dim=(0, 1)), y.double().std(dim=(0, 1), unbiased=False)
    assert torch.allclose(mean.reshape(-1), m64, rtol=1e-9, atol=1e-9) and torch.allclose(std.reshape(-1), s64, rtol=1e-8)
    mean2, std2 = fn.channel_mean_std(y)
    assert torch.equal(mean, mean2) and torch.equal(std, std2)
    # standardising with the measured statistics gives zero mean / unit variance per channel
    z = fn.noised_modality(x, seed=7, mean=mean, std=std)
    zm, zs = fn.channel_mean_std(z)
    assert float(zm.abs().max()) < 1e-4 and float((zs - 1).abs().max()) < 1e-4


@pytest.mark.gpu
def test_gpu_regenerated_tree_does_not_depend_on_the_launch_size(torch_cuda, tmp_path):
    """BASELINE configs[3] in small: the file tree from real rollouts, identical whichever way the worlds were packed into
    launches, with the device statistics equal to numpy's on the written file."""
    from conftest import blob_path
    rg, fn, ds = pkg("regenerate"), pkg("functions"), pkg("dataset")
    outs = []
    for wpl, sub in ((64, "a"), (24, "b")):
        roll = rg.DeviceRollouts({"softbox": blob_path("softbox")}, seed=5, worlds_per_launch=wpl)
        outs.append(rg.regenerate(str(tmp_path / sub), 40, 8, 8, roll, shapes=("softbox",), stats_fn=fn.channel_mean_std,
                                  drop_diverged=False, log=lambda *_: None, noise_seed=9))      # noise baked into */train
    assert [os.path.relpath(o["file"], str(tmp_path / "a")) for o in outs[0]] == [
        "sim_box/train.pickle", "sim_box/val.pickle", "sim_all/train.pickle", "sim_all/softbox_testing.pickle"]
    allk = []
    for a, b in zip(*outs):
        xa, ya = ds.read_pickle(a["file"])
        xb, yb = ds.read_pickle(b["file"])
        assert xa.shape == (a["samples"], 200, 12) and a["diverged"] == 0 and np.isfinite(xa).all()
        assert (ya == yb).all() and (xa == xb).all()
        st = np.load(a["file"].replace(".pickle", ".stats.npz"))
        assert np.allclose(st["mean"], xa.mean(axis=(0, 1)), rtol=1e-9, atol=1e-9) and np.allclose(st["std"], xa.std(axis=(0, 1)), rtol=1e-8)
        allk += list(ya)
    assert len(set(allk)) == len(allk) == 40 + 8 + 40 + 8          # no sample shared between files
    # the settle rows of a clean file are smooth, those of a noised */train file carry the accelerometer sigma
    xt, _ = ds.read_pickle(outs[0][0]["file"])
    xv, _ = ds.read_pickle(outs[0][1]["file"])
    assert 0.5 < np.diff(xt[:, 5:35, 0], axis=1).std() / np.sqrt(2) < 0.9 and np.diff(xv[:, 5:35, 0], axis=1).std() < 0.1


def test_emulated_trajectory_kernels_on_random_ragged_shapes(emulib):
    """Seeded sweep over odd shapes (worlds, rows per world and channel counts that are not multiples of the warp / CTA /
    grid sizes): all three kernels against their restatements."""
    rng = np.random.default_rng(1234)
    for case in range(12):
        W, T, nchan = int(rng.integers(1, 40)), int(rng.integers(1, 75)), int(rng.choice([4, 8, 12, 16, 24, 36, 64]))
        dtype = np.float32 if case % 2 == 0 else np.float64
        prec = 32 if dtype == np.float32 else 64
        x = _traj(W, T, nchan, dtype, seed=100 + case) + 500.0
        rows = W * T
        # statistics
        nbytes = emulib.sg_traj_stats_workspace_bytes(rows, nchan, 0)
        ws, mean, std = np.zeros(nbytes // 8), np.empty(nchan), np.empty(nchan)
        assert emulib.sg_traj_channel_stats(_vp(x), rows, nchan, prec, 0, _vp(mean), _vp(std), _vp(ws), nbytes, None) == 0, (W, T, nchan)
        wm, wsd = to.channel_mean_std(x)
        assert np.allclose(mean, wm.reshape(-1), rtol=1e-12) and np.allclose(std, wsd.reshape(-1), rtol=1e-9, atol=1e-10), (W, T, nchan)
        # noise with a row offset
        first = int(rng.integers(0, 1 << 20))
        nacc = int(rng.integers(0, nchan + 1))
        out = np.empty_like(x)
        assert emulib.sg_traj_add_noise(_vp(x), _vp(out), rows, first, nchan, nacc, 0.7, 0.06, case, None, None, prec, 0, None) == 0
        want = to.noised_modality(x, case, nacc=nacc, first_row=first)
        sig = np.where(np.arange(nchan) < nacc, 0.7, 0.06)
        assert (np.abs(out - want) <= sig * Z_TOL + np.abs(want) * (2e-7 if prec == 32 else 1e-15)).all(), (W, T, nchan)
        # mask, both modes
        touch = _touch(max(W, 2), T, seed=200 + case, p_finger=0.15)[:W]
        touch = np.ascontiguousarray(touch)
        for mode, m in (("intended", 0), ("reference", 1)):
            y = x.copy()
            left = np.full(W, 3, dtype=np.int32)
            assert emulib.sg_traj_mask_contact(_vp(y), _vp(touch), W, T, nchan, 3, 1 << 30, m, _vp(left) if m else None, prec, 0, None) == 0
            assert (y == to.mask_contact(x, touch, 2, 1 << 30, mode)).all(), (W, T, nchan, mode)


@pytest.mark.gpu
def test_gpu_trajectory_kernels_on_random_ragged_shapes(torch_cuda):
    """The ragged-shape sweep on the device (real grid sizes, ragged last CTAs / warps) through the Python mirror."""
    torch = torch_cuda
    fn = pkg("functions")
    rng = np.random.default_rng(4321)
    for case in range(10):
        W, T, nchan = int(rng.integers(1, 3000)), int(rng.integers(1, 75)), int(rng.choice([4, 8, 12, 16, 24, 36, 64]))
        dtype = np.float32 if case % 2 == 0 else np.float64
        x = _traj(W, T, nchan, dtype, seed=300 + case) + 500.0
        xd = torch.from_numpy(x).cuda()
        mean, std = fn.channel_mean_std(xd)
        wm, wsd = to.channel_mean_std(x)
        assert np.allclose(mean.cpu().numpy(), wm, rtol=1e-12) and np.allclose(std.cpu().numpy(), wsd, rtol=1e-9, atol=1e-10), (W, T, nchan)
        first, nacc = int(rng.integers(0, 1 << 20)), int(rng.integers(0, nchan + 1))
        out = fn.noised_modality(xd, seed=case, nacc=nacc, first_row=first)
        want = to.noised_modality(x, case, nacc=nacc, first_row=first)
        sig = np.where(np.arange(nchan) < nacc, 0.7, 0.06)
        assert (np.abs(out.cpu().numpy() - want) <= sig * Z_TOL + np.abs(want) * (2e-7 if dtype == np.float32 else 1e-15)).all(), (W, T, nchan)
