"""Parity of the CUDA path (through the C-ABI / BatchedManEnv / ManEnv) against the CPU oracle and the golden
vectors.  Bars: fp64 verification build <= 1e-9 relative per step (north_star asks 1e-5); fp32 fast path within
the stated tolerances below; discrete quantities (contact / row / sweep counts, touch masks) exact."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, blob_path, pkg

pytestmark = pytest.mark.gpu

# stated fp32 single-step tolerances, relative to the max magnitude of the quantity over the world
# about ten times what is measured on a B200 for softbox (profiles/r02x_test_gpu_measured.txt: q 1.1e-6, v 3.6e-5, qacc 1.6e-5,
# sens 1.0e-4 over the twelve golden snapshots); the larger models, in contact from the first step, get their own bound below
FP32_TOL = {"q": 1e-5, "v": 3e-4, "qacc": 2e-4, "sens": 1e-3}
FP32_TOL_OTHER = {"q": 5e-5, "v": 3e-3, "qacc": 3e-3, "sens": 1e-2}
FP64_TOL = 1e-10         # measured 8e-14 (softbox) .. 4e-13 (refined softbox); north_star asks 1e-5


def _note(text):
    """Measured values behind the asserted bounds, kept next to the run's other outputs (gpurun_out/ travels back)."""
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "test_gpu_measured.txt"), "a") as f:
            f.write(text + "\n")
    print(text)


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-12, np.abs(b).max()))


@pytest.fixture(scope="module")
def states():
    return np.load(os.path.join(GOLDEN, "softbox_states.npz"))


def make_env(batched, torch, name="softbox", W=4, dtype=None, **kw):
    return batched.BatchedManEnv(blob_path(name), W, dtype=dtype or torch.float64, **kw)


def one_step_from(env, q, v, act, warm, ctrl, W):
    env.set_state(q, v, act, warm)
    env.set_ctrl(np.tile(np.asarray(ctrl, dtype=np.float64), (W, 1)))
    sens, touch = env.step(1)
    return env.get_state(), sens.double().cpu().numpy(), env._touch.cpu().numpy()


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_single_step_parity_against_golden_states(torch_cuda, batched, states, prec):
    torch = torch_cuda
    W = 3
    env = make_env(batched, torch, W=W, dtype=torch.float64 if prec == "f64" else torch.float32)
    env.set_new_stiffness(stiffness=[700.0] * W)
    env.set_debug_world(1)
    worst = {}
    for i in range(len(states["step"])):
        (q1, v1, a1, qacc), sens, touch = one_step_from(env, states["q"][i], states["v"][i], states["act"][i], states["warm"][i],
                                                        [states["ctrl"][i]] * 2, W)
        assert int(env.debug(1, "ncon")[0]) == states["ncon1"][i]
        assert int(env.debug(1, "nefc")[0]) == states["nefc1"][i]
        assert int(touch[1]) == states["touch1"][i]
        err = {"q": rel(q1[1], states["q1"][i]), "v": rel(v1[1], states["v1"][i]), "qacc": rel(qacc[1], states["qacc1"][i]),
               "sens": rel(sens[1], states["sens1"][i])}
        for k, e in err.items():
            worst[k] = max(worst.get(k, 0.0), e)
            assert e <= (FP64_TOL if prec == "f64" else FP32_TOL[k]), (int(states["step"][i]), k, e)
        if prec == "f64":
            assert int(env.debug(1, "solver_iter")[0]) == states["iter1"][i]
            np.testing.assert_array_equal(q1[0], q1[2])          # worlds with identical inputs are bit-identical
        assert env.status()[1] == 0
    _note("single step %s vs golden states, worst over %d snapshots: %s" % (prec, len(states["step"]), ", ".join("%s %.2e" % kv for kv in sorted(worst.items()))))


def test_stage_diagnostics_match_oracle(torch_cuda, batched, states, make_world):
    """Contacts (dist/pos/frame, MuJoCo order) and constraint rows (aref, R, final forces) of one contact-rich step."""
    torch = torch_cuda
    i = list(states["step"]).index(850)
    env = make_env(batched, torch, W=1)
    env.set_new_stiffness(stiffness=[700.0])
    env.set_debug_world(0)
    one_step_from(env, states["q"][i], states["v"][i], states["act"][i], states["warm"][i], [states["ctrl"][i]] * 2, 1)
    w = make_world("softbox")
    w.set_state(states["q"][i], states["v"][i], states["act"][i], states["warm"][i]); w.set_ctrl([states["ctrl"][i]] * 2)
    w.step()
    for key, tol in (("con_dist", 1e-12), ("con_pos", 1e-12), ("con_frame", 1e-10), ("efc_aref", 1e-9), ("efc_R", 1e-11)):
        a, b = env.debug(0, key), w.get(key)
        assert a.shape == b.shape, key
        np.testing.assert_allclose(a, b, atol=tol * max(1.0, np.abs(b).max()), err_msg=key)
    f, fo = env.debug(0, "efc_force"), w.get("efc_force")
    np.testing.assert_allclose(f, fo, atol=1e-9 * np.abs(fo).max())


@pytest.mark.parametrize("k", [300.0, 700.0, 1400.0])
def test_full_horizon_rollout_fp64_vs_oracle(torch_cuda, batched, make_world, k):
    """Whole create_dataset episode on-chip vs the oracle: 200 x 12 sensor rows, touch masks, final state."""
    torch = torch_cuda
    env = make_env(batched, torch, W=2)
    traj, kk, st, touch = env.rollout(stiffness=[k, 700.0], return_touch=True)
    w = make_world("softbox", k=k)
    rows, otouch, ost = w.episode()
    assert ost == 0 and int(st[0]) == 0
    g = traj[0].cpu().numpy()
    scale = np.abs(rows).max(axis=0)
    err = np.abs(g - rows) / scale                    # (200, 12), relative to each channel's peak
    # until the hand closes there is no chaos to amplify round-off: agreement at the 1e-9 level
    assert err[:40].max() < 1e-9, err[:40].max()
    # over the contact-rich squeeze the 1e-13 per-step differences (FMA contraction, reduction order) are
    # amplified by make/break events; the drift stays bounded and most rows still agree tightly
    row_err = err.max(axis=1)
    print("k=%g full-horizon drift: max %.2e  median %.2e  rows>1e-5: %d" % (k, row_err.max(), np.median(row_err), int((row_err > 1e-5).sum())))
    _note("fp64 full horizon k=%g: row max %.3e median %.3e rows>1e-5 %d" % (k, row_err.max(), np.median(row_err), int((row_err > 1e-5).sum())))
    # Bound = the oracle's OWN sensitivity at this stiffness: the same episode re-run on the host with the stiffness moved by
    # +-1e-12 and +-3e-13 relative (what a different summation order does in fp64).  Where the oracle is well conditioned
    # (k = 300, 1000: worst row 2e-6) the kernel has to agree that tightly; where a make/break event bifurcates (k = 1400:
    # the -1e-12 run moves the oracle's own worst row by 4e-2 and its median row by 3e-4) no fp64 implementation can do
    # better than that envelope.  Rows before the oracle's first sensitive row must agree to 1e-5 regardless.
    env_max, env_med, first, env_q = 0.0, 0.0, rows.shape[0], 0.0
    oq = w.get_state()[0]
    for eps in (1e-12, -1e-12, 3e-13, -3e-13):
        w2 = make_world("softbox", k=k * (1.0 + eps))
        e2 = (np.abs(w2.episode()[0] - rows) / scale).max(axis=1)
        env_q = max(env_q, rel(w2.get_state()[0], oq))
        env_max = max(env_max, float(e2.max())); env_med = max(env_med, float(np.median(e2)))
        if (e2 > 1e-6).any(): first = min(first, int(np.argmax(e2 > 1e-6)))
    _note("fp64 full horizon k=%g: oracle self-sensitivity envelope (+-1e-12): row max %.3e median %.3e first sensitive row %d" % (k, env_max, env_med, first))
    assert row_err[:max(40, first - 5)].max() < 1e-5
    assert row_err.max() < max(1e-5, 3.0 * env_max) and np.median(row_err) < max(1e-8, 3.0 * env_med), (row_err.max(), env_max, np.median(row_err), env_med)
    assert row_err.max() < 0.2
    tg = touch[0].cpu().numpy()
    assert (tg != otouch).sum() <= 4
    q, v, a, qacc = env.get_state()
    oq, ov, oa, _ = w.get_state()
    # the final state sits after ~700 contact-rich steps of chaotic amplification: the filter state `act` is
    # contact-free and must agree tightly; qpos is held to a drift bound (median over dofs), not to parity
    assert rel(a[0], oa) < 1e-9
    _note("fp64 full horizon k=%g: final qpos median %.3e max %.3e" % (k, float(np.median(np.abs(q[0] - oq))) / np.abs(oq).max(), rel(q[0], oq)))
    assert float(np.median(np.abs(q[0] - oq))) / np.abs(oq).max() < 1e-5 and rel(q[0], oq) < max(1e-4, 3.0 * env_q)
    if k == 700.0:
        gold = np.load(os.path.join(GOLDEN, "softbox_episode_k700.npz"))
        assert (np.abs(g - gold["rows"]) / scale)[:40].max() < 1e-9
        np.testing.assert_array_equal(traj[1].cpu().numpy(), g)   # same stiffness, other CTA: identical


def test_step_api_equals_rollout(torch_cuda, batched):
    """ManEnv-style stepping (state round-trips through HBM every env-step) == one-launch rollout, bit for bit."""
    torch = torch_cuda
    sched = batched.default_schedule(2, n_settle=3, n_iter=12, open_close_div=6)
    env = make_env(batched, torch, W=2)
    traj, k, st = env.rollout(schedule=sched, stiffness=[500.0, 900.0])
    env2 = make_env(batched, torch, W=2)
    env2.reset(stiffness=[500.0, 900.0])
    rows = []
    ev, val = sched
    for t in range(ev.shape[0]):
        if ev[t]:
            env2.set_ctrl(np.tile(val[t], (2, 1)))
        r, c = env2.step()
        rows.append(r)
    np.testing.assert_array_equal(torch.stack(rows, 1).cpu().numpy(), traj.cpu().numpy())


def test_fp32_drift_and_feature_statistics(torch_cuda, batched):
    """fp32 fast path over the full horizon: finite, nothing diverges, and the per-stiffness-bin feature
    statistics (|acc|, |gyro| mean/peak) agree with the fp64 build (north_star: distributions match)."""
    torch = torch_cuda
    ds = pkg("dataset")
    W = 192
    e32 = make_env(batched, torch, W=W, dtype=torch.float32, seed=5)
    e64 = make_env(batched, torch, W=W, dtype=torch.float64, seed=5)
    t32, k32, s32 = e32.rollout()
    t64, k64, s64 = e64.rollout()
    assert torch.isfinite(t32).all() and int((s32 != 0).sum()) == 0 and int((s64 != 0).sum()) == 0
    np.testing.assert_array_equal(k32.cpu().numpy(), k64.cpu().numpy())
    a, b = t32.double().cpu().numpy(), t64.cpu().numpy()
    # before the hand closes the trajectories coincide closely; afterwards contact chaos decorrelates details
    pre = np.abs(a[:, :40] - b[:, :40]).max() / np.abs(b[:, :40]).max()
    assert pre < 1e-4, pre
    med = np.median(np.abs(a - b), axis=(0, 1)) / np.abs(b).max(axis=(0, 1))
    assert med.max() < 5e-3, med
    _, sa = ds.feature_stats(a, k32.cpu().numpy(), nbins=3)
    _, sb = ds.feature_stats(b, k64.cpu().numpy(), nbins=3)
    for x, y in zip(sa, sb):
        np.testing.assert_allclose(x["mean"], y["mean"], rtol=0.05)
        np.testing.assert_allclose(x["std"], y["std"], rtol=0.15)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_trace_error_statistic_of_the_bench_against_the_oracle(torch_cuda, batched, prec):
    """The metric's second half, asserted: the statistic bench.py prints as `sensor_trace_error` (CUDA path vs the fp64
    ORACLE with the same per-world stiffness / damping / object offset of BASELINE configs[2], relative to each channel's
    peak) stays inside bench.TRACE_TOLERANCE.  The horizon bound is set by the dynamics, not by the arithmetic: the
    oracle run against itself from a 1e-6 relative parameter perturbation deviates by a median of up to 3e-2
    (tests/test_oracle.py::test_oracle_sensitivity_bounds_the_horizon_tolerances)."""
    torch = torch_cuda
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    args = argparse.Namespace(randomise="all", fixed_stiffness=None, seed=11, damping_range=[100.0, 200.0], offset_range=0.05,
                              tendon_damping=None)
    W = 12
    ids = np.arange(1000, 1000 + W)
    k, d, off = bench.world_params(args, ids, 0)
    env = make_env(batched, torch, W=W, dtype=torch.float32 if prec == "f32" else torch.float64)
    env.set_params(damping=d, object_offset=off)
    traj, _, st = env.rollout(stiffness=k)
    assert int((st != 0).sum()) == 0
    out = bench.sensor_trace_error(args, blob_path("softbox"), traj.double().cpu().numpy(), (k, d, off))
    _note("trace error %s: %s" % (prec, {kk: vv for kk, vv in out.items() if isinstance(vv, float)}))
    tol = bench.TRACE_TOLERANCE["fp32" if prec == "f32" else "fp64"]
    assert "error" not in out, out
    assert out["settle_rows_max_rel"] <= tol["settle_rows_max_rel"], out
    assert out["row_median_rel"] <= tol["row_median_rel"], out
    assert out["row_median_rel_worst_world"] <= 4 * tol["row_median_rel"], out


def test_rollout_soa_layout_equals_world_major(torch_cuda, batched):
    """north-star item (5): the SoA trajectory [T,12,W] (world index fastest) holds the same values as the world-major
    [W,T,12] sample layout, bit for bit, in both precisions."""
    torch = torch_cuda
    sched = batched.default_schedule(2, n_settle=2, n_iter=10, open_close_div=5)
    for dtype in (torch.float32, torch.float64):
        env = make_env(batched, torch, W=70, dtype=dtype)
        k = np.linspace(300, 1400, 70)
        a, _, _ = env.rollout(schedule=sched, stiffness=k)
        b, _, _ = env.rollout(schedule=sched, stiffness=k, layout="TCW")
        assert tuple(b.shape) == (12, 12, 70) and tuple(a.shape) == (70, 12, 12)
        np.testing.assert_array_equal(b.permute(2, 0, 1).cpu().numpy(), a.cpu().numpy())


def test_per_world_parameters_parity(torch_cuda, batched, states, make_world):
    """stiffness / shell damping / tendon damping / object offset overrides against the oracle with the same edits."""
    torch = torch_cuda
    i = list(states["step"]).index(450)
    W = 2
    env = make_env(batched, torch, W=W)
    env.set_new_stiffness(stiffness=[400.0, 1250.0])
    env.set_params(damping=[80.0, 150.0], tendon_damping=[60.0, 100.0], object_offset=[[0.01, -0.02, 0.03], [0.0, 0.0, 0.0]])
    (q1, v1, a1, qacc), sens, touch = one_step_from(env, states["q"][i], states["v"][i], states["act"][i], states["warm"][i],
                                                    [states["ctrl"][i]] * 2, W)
    for wi, (k, d, td, off) in enumerate(((400.0, 80.0, 60.0, [0.01, -0.02, 0.03]), (1250.0, 150.0, 100.0, [0, 0, 0]))):
        w = make_world("softbox", k=k)
        for dof in range(8, 118):
            w.set_dof_damping(dof, d)
        w.set_tendon_damping(0, td)
        w.set_body_pos(10, np.array([1.7, 0, 1.0]) + off)
        w.set_state(states["q"][i], states["v"][i], states["act"][i], states["warm"][i]); w.set_ctrl([states["ctrl"][i]] * 2)
        w.step()
        oq, ov, _, oacc = w.get_state()
        assert rel(q1[wi], oq) < FP64_TOL and rel(v1[wi], ov) < FP64_TOL and rel(qacc[wi], oacc) < FP64_TOL
        assert rel(sens[wi], w.sensordata()) < FP64_TOL


@pytest.mark.parametrize("name", ["softball", "softcylinder"])
def test_other_models_first_steps(torch_cuda, batched, make_world, name):
    """Ball / cylinder (45 / 37 initial penetrations, elliptic contacts from step 0): per-step parity while the
    oracle itself is still finite (these models run away under the restated semantics)."""
    torch = torch_cuda
    env = make_env(batched, torch, name=name, W=1)
    env.set_new_stiffness(stiffness=[700.0])
    env.set_debug_world(0)
    w = make_world(name)
    w.reset()
    for step in range(12):
        q, v, a, ws = w.get_state()
        if w.step():
            break
        (q1, v1, a1, qacc), sens, touch = one_step_from(env, q, v, a, ws, [0, 0], 1)
        oq, ov, _, oacc = w.get_state()
        assert int(env.debug(0, "ncon")[0]) == w.get_int("ncon")
        assert rel(q1[0], oq) < 1e-8 and rel(v1[0], ov) < 1e-8 and rel(qacc[0], oacc) < 1e-8, step
    assert step >= 5


def test_divergence_is_flagged_and_reset(torch_cuda, batched):
    torch = torch_cuda
    env = make_env(batched, torch, W=3, dtype=torch.float32)
    q = np.zeros((3, 118)); q[1, 30] = 1e11
    env.set_state(q, np.zeros((3, 118)), np.zeros((3, 2)), np.zeros((3, 118)))
    env.step(2)
    st = env.status(clear=True)
    assert st[1] & batched.ST_DIVERGED and st[0] == 0 and st[2] == 0
    assert np.abs(env.get_state()[0][1]).max() < 1e-3
    assert (env.status() == 0).all()


def test_contact_modes(torch_cuda, batched):
    torch = torch_cuda
    sched = batched.default_schedule(2, n_settle=2, n_iter=30, open_close_div=80)
    flags = {}
    for mode in ("intended", "reference"):
        env = make_env(batched, torch, W=1, contact_mode=mode)
        env.reset(stiffness=[700.0])
        out = []
        for t in range(32):
            if t == 2:
                env.close_hand()
            if t == 20:
                env.loose_hand()
            r, c = env.step()
            out.append(bool(c[0]))
        flags[mode] = out
    assert not any(flags["intended"][:3]) and any(flags["intended"])
    first = flags["intended"].index(True)
    assert flags["reference"][:first] == [False] * first and flags["reference"][first]


def test_sharding_invariance(torch_cuda, batched):
    """World w draws the same stiffness and produces the same trace whichever shard it lands in."""
    torch = torch_cuda
    sched = batched.default_schedule(2, n_settle=1, n_iter=4, open_close_div=80)
    full = make_env(batched, torch, W=6, dtype=torch.float32, seed=3)
    tf, kf, _ = full.rollout(schedule=sched)
    part = make_env(batched, torch, W=3, dtype=torch.float32, seed=3, world_offset=3)
    tp, kp, _ = part.rollout(schedule=sched)
    np.testing.assert_array_equal(kf[3:].cpu().numpy(), kp.cpu().numpy())
    np.testing.assert_array_equal(tf[3:].cpu().numpy(), tp.cpu().numpy())


def test_host_buffer_entry_point(torch_cuda, batched):
    """sg_batch_rollout_host (the e2e path bench.py times) returns the same trajectory as the device path."""
    import ctypes as C
    torch = torch_cuda
    lib = pkg("_lib")
    sched = batched.default_schedule(2, n_settle=1, n_iter=5, open_close_div=80)
    env = make_env(batched, torch, W=5, dtype=torch.float32)
    k = np.linspace(300, 1400, 5)
    traj, _, _ = env.rollout(schedule=sched, stiffness=k)
    ev, val = sched
    sc = lib.SgSchedule(1, 7, ev.shape[0], ev.ctypes.data_as(C.POINTER(C.c_int)), val.ctypes.data_as(C.POINTER(C.c_double)))
    out = np.zeros((5, ev.shape[0], 12), dtype=np.float32)
    status = np.ones(5, dtype=np.int32)
    lib.check(env.L.sg_batch_rollout_host(env.h, C.byref(sc), k.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), None,
                                          status.ctypes.data_as(C.c_void_p)))
    np.testing.assert_array_equal(out, traj.cpu().numpy())
    assert (status == 0).all()


def test_manenv_dropin_episode(torch_cuda, make_world):
    """The reference driver's episode through the drop-in ManEnv (one world, fp64) == the oracle's episode."""
    sys.path.insert(0, os.path.join(ROOT, "soft-grip_b200"))
    from environment import ManEnv
    cd = pkg("create_dataset")
    ManEnv.finger_names = ['g12', 'g2']                      # undo the class-level consumption of earlier tests
    np.random.seed(11)
    env = ManEnv(sim_start=1, sim_step=7, env_paths=[blob_path("softbox")], is_vis=False)
    k = env.reset()
    rows = np.array(list(cd.episode_rows(env, mask_contact=False)))
    assert rows.shape == (200, 12) and 300 <= k <= 1400
    w = make_world("softbox", k=k)
    orows, otouch, st = w.episode()
    scale = np.abs(orows).max(axis=0)
    assert (np.abs(rows - orows).max(axis=0) / scale).max() < 1e-5
    out = env.step()
    assert out.dtype == object and out.shape == (2,) and out[0].shape == (12,)
    readings, contact = out
    assert isinstance(contact, bool)
    assert ManEnv.finger_names == []                         # reference quirk: the class list was consumed
    sim = env.get_env()
    assert sim.model.jnt_stiffness[11] == k and sim.model.jnt_stiffness[10] == 700 and sim.model.tendon_stiffness[0] == k
    ManEnv.finger_names = ['g12', 'g2']


def test_many_worlds_identical_inputs_are_identical(torch_cuda, batched):
    """BASELINE.json configs[1] shape (4096 worlds, fixed stiffness): every CTA computes the same bits."""
    torch = torch_cuda
    sched = batched.default_schedule(2, n_settle=2, n_iter=10, open_close_div=5)
    env = make_env(batched, torch, W=4096, dtype=torch.float32)
    traj, k, st = env.rollout(schedule=sched, stiffness=np.full(4096, 700.0))
    assert int((st != 0).sum()) == 0
    assert bool((traj == traj[0:1]).all())


def test_contact_capacity_overflow_keeps_the_first_contacts(torch_cuda, batched, states, mjcf, oracle, tmp_path):
    """More active contacts than the per-world capacity (MuJoCo: nconmax warning): the first `maxcon` contacts in MuJoCo's
    order keep their rows, the rest are dropped and SG_ST_CON_FULL is set -- the same step as the oracle at that nconmax."""
    torch = torch_cuda
    cap, W = 16, 3
    model = mjcf.load_blob(blob_path("softbox"))
    model.opt["nconmax"] = cap
    blob = mjcf.model_to_blob(model)
    p = tmp_path / "softbox_cap.sgm"
    p.write_bytes(blob)
    ow = oracle.OracleWorld(oracle.OracleModel(blob))
    ow.set_geom_mask(batched.geom_name_mask(model.names["geom"], "OBJ", ("g12", "g2")))
    ow.set_stiffness(700.0)
    os.environ["SOFTGRIP_MAXCON"] = str(cap)
    try:
        env = batched.BatchedManEnv(str(p), W, dtype=torch.float64)
    finally:
        os.environ.pop("SOFTGRIP_MAXCON", None)
    env.set_new_stiffness(stiffness=[700.0] * W)
    env.set_debug_world(1)
    for i in (4, 6, 7):                               # 8 (fits), 35 and 58 contacts
        env.status(clear=True)
        (q1, v1, a1, qacc), sens, touch = one_step_from(env, states["q"][i], states["v"][i], states["act"][i], states["warm"][i],
                                                        [states["ctrl"][i]] * 2, W)
        ow.set_state(states["q"][i], states["v"][i], states["act"][i], states["warm"][i])
        ow.set_ctrl([states["ctrl"][i]] * 2)
        ost = ow.step()
        oq, ov, oa, owarm = ow.get_state()
        full = states["ncon1"][i] > cap
        assert int(env.debug(1, "ncon_rows")[0]) == min(cap, int(states["ncon1"][i])) == ow.get_int("ncon")
        assert bool(ost & 2) == full and all(bool(s & 2) == full for s in env.status())
        assert rel(q1[1], oq) < FP64_TOL and rel(v1[1], ov) < FP64_TOL and rel(qacc[1], owarm) < FP64_TOL


# under the restated semantics the volume mode is stable iff (1 + n d_t / d_j) / 65.7 < 2, i.e. n d_t < ~130 d_j:
# ball / cylinder (n = 218 / 192) need a volume-tendon damper d_t <= ~70, the refined softbox of BASELINE configs[4]
# (7x9x13 grid, n = 434) d_t <= ~30; the committed files have d_t = d_j = 100.  These are the stated stable values.
STABLE_TENDON_DAMPING = {"softball": 50.0, "softcylinder": 50.0, "softbox_refined": 20.0}


def _protocol_ctrl(step):
    """ctrl of create_dataset's episode at physics step `step` (ref: create_dataset.py:33-60 with sim_start 1, sim_step 7)."""
    return 0.0 if step < 281 else (-0.2 if step < 841 else 0.2)


@pytest.mark.parametrize("name", ["softball", "softcylinder", "softbox_refined"])
def test_other_models_along_a_stabilised_episode(torch_cuda, batched, make_world, name):
    """SURVEY 8d cfg 3: with the volume-tendon damper at a stable value (stated above) ball, cylinder and the refined softbox of configs[4] run the whole squeeze
    episode clean; fp64 step parity from oracle snapshots all along it, and the fp64 / fp32 rollouts finish with no world
    flagged (drift against the oracle is printed, not asserted: contacts exist from step 0, so round-off is amplified
    from the first row on)."""
    torch = torch_cuda
    w = make_world(name)
    w.set_tendon_damping(0, STABLE_TENDON_DAMPING[name])
    w.reset()
    snaps, want = {}, (0, 150, 285, 500, 838, 1000, 1395)
    for step in range(1401):
        w.set_ctrl([_protocol_ctrl(step)] * 2)
        if step in want:
            snaps[step] = w.get_state()
        assert w.step() == 0, step
        if step in want:
            snaps[step] = (snaps[step], w.get_state(), w.get_int("ncon"))
    env = make_env(batched, torch, name=name, W=2)
    env.set_new_stiffness(stiffness=[700.0, 700.0])
    env.set_params(tendon_damping=[STABLE_TENDON_DAMPING[name]] * 2)
    env.set_debug_world(1)
    for step in want:
        (q, v, a, ws), (oq, ov, oa, oacc), ncon = snaps[step]
        (q1, v1, a1, qacc), sens, touch = one_step_from(env, q, v, a, ws, [_protocol_ctrl(step)] * 2, 2)
        assert int(env.debug(1, "ncon")[0]) == ncon
        assert rel(q1[1], oq) < 1e-8 and rel(v1[1], ov) < 1e-8 and rel(qacc[1], oacc) < 1e-8, step
    assert (env.status() == 0).all()
    w2 = make_world(name)
    w2.set_tendon_damping(0, STABLE_TENDON_DAMPING[name])
    rows, otouch, ost = w2.episode()
    assert ost == 0
    for dtype in (torch.float64, torch.float32):
        e = make_env(batched, torch, name=name, W=64, dtype=dtype, seed=2)
        e.set_params(tendon_damping=[STABLE_TENDON_DAMPING[name]] * 64)
        traj, k, st = e.rollout(stiffness=[700.0] + list(np.linspace(300, 1400, 63)))
        flagged = st if dtype == torch.float64 else (st & batched.ST_DIVERGED)      # fp32: no world may run away
        assert bool(torch.isfinite(traj).all()) and int((flagged != 0).sum()) == 0, (name, dtype, st.cpu().numpy())
        err = (np.abs(traj[0].double().cpu().numpy() - rows) / np.abs(rows).max(axis=0)).max(axis=1)
        print("%s %s stabilised episode vs oracle: first row %.2e, median row %.2e, max row %.2e" % (name, dtype, err[0], np.median(err), err.max()))


def test_plain_c_caller_matches_the_python_path(torch_cuda, batched, tmp_path):
    """tests/cabi/rollout_host.c -- a C99 program with host buffers only -- produces the same bits as BatchedManEnv.rollout."""
    import subprocess
    torch = torch_cuda
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_abi import build_c_caller
    lib = pkg("_lib")
    exe = build_c_caller(lib.LIB_PATH, tmp_path)
    W = 6
    out = subprocess.run([exe, blob_path("softbox"), str(W), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    raw = np.fromfile(str(tmp_path / "out.bin"), dtype=np.float32, count=W * 200 * 12).reshape(W, 200, 12)
    status = np.fromfile(str(tmp_path / "out.bin"), dtype=np.int32, offset=W * 200 * 12 * 4)
    assert status.shape == (W,) and (status == 0).all()
    env = make_env(batched, torch, W=W, dtype=torch.float32)
    traj, k, st = env.rollout(stiffness=[300.0 + 1100.0 * w / (W - 1) for w in range(W)])
    np.testing.assert_array_equal(traj.cpu().numpy(), raw)


def test_tensor_memory_rows_and_record_ring_change_no_bit(torch_cuda, batched, monkeypatch):
    """The default geometry (16 warps, one CTA per SM) keeps the equality rows in tensor memory and feeds the contact
    blocks from the shared-memory record ring (cp.async one block ahead).  Both are storage changes: the trajectories of
    a squeeze episode with per-world stiffness, damping and pose equal those of SOFTGRIP_TMEM=0 (rows in shared memory,
    records read from the scratch) and of SOFTGRIP_RING=0 bit for bit, in both precisions; a small batch (a few warps
    per CTA, tensor memory forced) as well."""
    torch = torch_cuda
    sched = batched.default_schedule(2, n_settle=10, n_iter=40, open_close_div=20)
    for dtype, W in ((torch.float32, 9472 + 37), (torch.float64, 700), (torch.float32, 19)):
        rng = np.random.default_rng(5)
        k, d, off = rng.uniform(300, 1400, W), rng.uniform(100.0, 200.0, W), rng.uniform(-0.05, 0.05, (W, 3))
        out = []
        for tmem, ring in ((None if W > 100 else "1", None), ("0", None), (None if W > 100 else "1", "0")):
            for key, val in (("SOFTGRIP_TMEM", tmem), ("SOFTGRIP_RING", ring)):
                if val is None: monkeypatch.delenv(key, raising=False)
                else: monkeypatch.setenv(key, val)
            env = make_env(batched, torch, W=W, dtype=dtype)
            tm = env.debug(0, "tensor_memory")
            assert (tm[0] > 0) == (tmem != "0") and (tm[2] > 0) == (tmem != "0" and ring != "0"), (tmem, ring, tm)
            env.set_params(damping=d, object_offset=off)
            traj, _, st = env.rollout(schedule=sched, stiffness=k)
            assert int((st.cpu().numpy() & 1).sum()) == 0
            out.append(traj.cpu().numpy())
            del env
        np.testing.assert_array_equal(out[0], out[1])
        np.testing.assert_array_equal(out[0], out[2])
        assert np.abs(out[0][:, -1, :]).max() > 0
