#!/usr/bin/env python
"""bench.py — the SDF-contact hot path on the C2 workload (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E]

One "step" = one pass of the hot path over one batch: for every environment of a
1024-rollout two-part BFS planning batch, detection of every contact point of
both pairs against the other part's SDF (LOOP C), penalty force + Jacobian
evaluation of every active contact (LOOP D) and the accumulation into the
per-environment system arrays fm / Km / Dm — i.e. what Robot::computeForceWith
Derivative's loop over the contact forces does (reference Robot.cpp:546-547),
for all environments at once.

metric  contact_point_sdf_queries_per_sec = detection queries processed by the
        whole pass (E * sum_p N_p) / step time.  `value`: inputs resident in HBM,
        CUDA events. `e2e`: the same pass through the host-buffer C-ABI call
        (sdfc_batch_eval), H2D of poses and D2H of fm/Km/Dm inside the timing.
--impl reference  times the reference's own CPU implementation (oracle/_ref, the
        unmodified RedMax objects; falls back to the C oracle port) on all host
        cores over a bounded sample of the same environments.
Multi-GPU: environments are independent, so every rank takes its own E rollouts
(weak scaling) with no data-path collective; one all_gather of the per-rank
outcome record (contact counts) at the end, as the planner's final gather.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contact_point_sdf_queries_per_sec"
UNIT = "queries/s"
# SURVEY.md §8(d) unit per detection query: 24 B point + eight-corner gather + 8 B result. The gather is 64 B for an
# fp64 grid and 32 B when the grid is stored fp32 (legal for freshly generated grids): 96 B / 64 B. bench reports the
# unit of the storage it actually uses and the 96 B figure beside it.
BYTES_PER_QUERY_FP64 = 96
BYTES_PER_QUERY_FP32 = 64
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
# FP64 flops per active contact of the planner kernel (contact_eval_nf + warp reduction), FMA = 2: hand census of
# the source; profiles/flop_census.json holds the executed-instruction count measured with ncu.
FLOPS_PER_CONTACT_NF_CENSUS = 2307


def workload(n_envs, seed_offset=0):
    from assemble_them_all_b200 import scenes, synth
    return scenes.c2_bfs_rollouts(n_envs=n_envs, n_vertices=10000, seed=synth.MESH_SEED + seed_offset)


# ---------------------------------------------------------------------------
# reference CPU arm (oracle/_ref = unmodified reference objects; else the C port)
# ---------------------------------------------------------------------------
_W = {}


def _ref_worker_init(scene_blob, use_ref):
    """Per process: build the checker once and inject the scene's arrays."""
    sc = scene_blob
    _W["scene"] = sc
    _W["kind"] = "port"
    if use_ref:
        try:
            from oracle import ref_scene
            _W["probe"] = ref_scene.make_probe(sc["bodies"], sc["kn"], sc["kt"], sc["mu"], sc["damping"],
                                               joint_type=sc.get("joint_type", "translational"),
                                               inject_frames=sc.get("inject_frames", False))
            _W["kind"] = "reference"
            return
        except Exception as exc:  # pragma: no cover - only when oracle/_ref is absent
            _W["ref_error"] = repr(exc)
    from oracle import sdfc_oracle as O
    _W["O"] = O
    _W["ogrids"] = [O.Grid.from_dict(g) for g in sc["grids"]]


def _ref_rollout_envs(env_ids):
    """Simulation::forward(frame_skip) of the given rollouts on this core; -> (seconds, env-steps, kind, finals)."""
    sc = _W["scene"]
    n = sc["frame_skip"]
    t_total, finals = 0.0, []
    if _W["kind"] == "reference":
        from oracle import ref_scene
        for e in env_ids:
            t, q, qd = ref_scene.rollout(_W["probe"], sc["q0"][e], sc["qdot0"][e], sc["f_ext"][e, 0], n)
            t_total += t
            finals.append((e, q, qd))
    else:
        from oracle import sim_oracle as S
        jt = {"fixed": S.FIXED, "translational": S.TRANSLATIONAL, "free3d-exp": S.FREE3D_EXP}
        for e in env_ids:
            bodies = [S.BodyDesc(jt[b.joint_type], np.eye(4), b.rest_E, b.inertia, b.mass) for b in sc["bodies"]]
            pairs = [S.PairDesc(a, b, _W["ogrids"][b], sc["points"][a], sc["kn"], sc["kt"], sc["mu"], sc["damping"])
                     for a, b in sc["pairs"]]
            sim = S.Sim(bodies, pairs, sc["h"], sc["gravity"])
            sim.q, sim.qdot, sim.f_ext = sc["q0"][e].copy(), sc["qdot0"][e].copy(), sc["f_ext"][e].copy()
            t0 = time.perf_counter()
            for _ in range(n):
                sim.step()
            t_total += time.perf_counter() - t0
            finals.append((e, sim.q, sim.qdot))
    return t_total, len(env_ids) * n, _W["kind"], finals


def _ref_eval_envs(args):
    """Evaluate (with derivatives) the given environments once; -> (seconds, n_envs, kind, checksum)."""
    env_ids, want_outputs = args
    sc = _W["scene"]
    outs = []
    t_total = 0.0
    if _W["kind"] == "reference":
        p = _W["probe"]
        for e in env_ids:
            for b, name in enumerate(("part0", "part1")):
                p.set_body_pose(name, sc["E"][e, b], sc["phi"][e, b])
            t_total += p.time_force_evals(1, True)
            if want_outputs:
                acc = [p.force_deriv_public(fi) for fi in range(p.num_forces())]
                outs.append((sum(a[0] for a in acc), sum(a[1] for a in acc), sum(a[2] for a in acc)))
    else:
        O = _W["O"]
        for e in env_ids:
            t0 = time.perf_counter()
            fs, Ks, Ds = np.zeros(12), np.zeros((12, 12)), np.zeros((12, 12))
            for (a, b) in sc["pairs"]:
                fm, Km, Dm, _ = O.pair_force(_W["ogrids"][b], sc["E"][e, a], sc["phi"][e, a], sc["E"][e, b], sc["phi"][e, b],
                                             sc["points"][a], sc["kn"], sc["kt"], sc["mu"], sc["damping"], True)
                O.scatter(12, 6 * a, 6 * b, fm, Km, Dm, fs, Ks, Ds)
            t_total += time.perf_counter() - t0
            if want_outputs:
                outs.append((fs, Ks, Ds))
    return t_total, len(env_ids), _W["kind"], outs


def scene_blob(scene):
    def gd(g):
        return {"ni": g.ni, "nj": g.nj, "nk": g.nk, "min_box": g.min_box, "max_box": g.max_box, "dx": g.dx, "dy": g.dy,
                "dz": g.dz, "col_th": g.col_th, "grid": g.grid}
    blob = {"points": [b.points for b in scene.bodies], "grids": [gd(b.grid) for b in scene.bodies],
            "bodies": scene.bodies, "pairs": list(scene.pairs), "kn": scene.kn, "kt": scene.kt, "mu": scene.mu,
            "damping": scene.damping, "col_th": scene.bodies[0].grid.col_th}
    if hasattr(scene, "q0"):   # RolloutScene (stepper)
        blob.update({"q0": scene.q0, "qdot0": scene.qdot0, "f_ext": scene.f_ext, "h": scene.h, "gravity": scene.gravity,
                     "frame_skip": scene.frame_skip, "joint_type": scene.bodies[0].joint_type, "inject_frames": True})
    else:
        blob.update({"E": scene.E, "phi": scene.phi})
    return blob


def have_ref():
    from oracle import ref_scene
    return ref_scene.available()


def rollout_workload(n_envs, seed_offset=0):
    from assemble_them_all_b200 import scenes, synth
    return scenes.c2_planner_rollouts(n_envs=n_envs, n_vertices=10000, seed=synth.MESH_SEED + seed_offset)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    scene = workload(args.envs)
    q_per_env = scene.queries_per_eval // scene.n_envs
    sample = min(scene.n_envs, max(cores, args.ref_envs_per_core * cores))
    blob = scene_blob(scene)
    if have_ref():
        # the workers are forked and terminated; the parent maps the reference module as well, so that the
        # driver's record of loaded native libraries shows which CPU implementation this arm ran
        from oracle import ref_scene
        ref_scene._import_probe()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_ref_worker_init, initargs=(blob, have_ref())) as pool:
        shards = [list(range(i, sample, cores)) for i in range(cores)]
        shards = [(s, False) for s in shards if s]
        kind = "port"
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_eval_envs, shards)
            dt = time.perf_counter() - t0
            kind = res[0][2]
            if it >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample * q_per_env / (ms * 1e-3)
    # env-steps/s of the reference's own Simulation::forward on the planner rollouts: one rollout
    # (frame_skip BDF1 steps) per core, measured once
    rs = rollout_workload(max(cores, 16))
    with ctx.Pool(cores, initializer=_ref_worker_init, initargs=(scene_blob(rs), have_ref())) as pool:
        rres = pool.map(_ref_rollout_envs, [[e] for e in range(cores)])
        dt_roll = max(r[0] for r in rres)  # all cores run concurrently: time of the slowest rollout (setup excluded)
    env_steps = {"value": sum(r[1] for r in rres) / dt_roll, "unit": "env-steps/s", "cores": cores, "kind": rres[0][2],
                 "sample": f"{cores} rollouts x {rs.frame_skip} BDF1 steps, one process per core concurrently, "
                           f"slowest rollout {dt_roll:.2f} s in Simulation::forward",
                 "workload": rs.label}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": scene.label, "envs_per_step": sample, "queries_per_env": q_per_env, "derivatives": True},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} of {scene.n_envs} environments per step, one process per core (fork), "
                                   f"Robot-level loop over both contact forces with derivatives"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "env_steps": env_steps,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for i, n in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def cpu_baseline_leg(scene, gpu_out, budget_s=20.0):
    """Rank 0, N=1: time the reference CPU path on one core over a bounded sample of
    the same environments and gate parity of the GPU results on that sample."""
    blob = scene_blob(scene)
    _ref_worker_init(blob, have_ref())
    q_per_env = scene.queries_per_eval // scene.n_envs
    fm_g, Km_g, Dm_g = gpu_out
    t_total, n_done, worst = 0.0, 0, 0.0
    t_start = time.perf_counter()
    order = list(range(0, scene.n_envs, max(1, scene.n_envs // 97))) + list(range(scene.n_envs))
    seen = set()
    for e in order:
        if e in seen:
            continue
        seen.add(e)
        check = n_done < 16
        t, _, kind, outs = _ref_eval_envs(([e], check))
        t_total += t
        n_done += 1
        if check:
            fm, Km, Dm = outs[0]
            for ref, got in ((fm, fm_g[e]), (Km, Km_g[e]), (Dm, Dm_g[e])):
                ref = np.asarray(ref)
                s = np.max(np.abs(ref))
                err = float(np.max(np.abs(ref - got)) / s) if s > 0 else float(np.max(np.abs(got)))
                worst = max(worst, err)
        if time.perf_counter() - t_start > budget_s:
            break
    if worst > 1e-9:
        raise SystemExit(f"PARITY FAILURE: GPU vs CPU {_W['kind']} relative error {worst:.3e} > 1e-9")
    return {"value": n_done * q_per_env / t_total, "unit": UNIT, "cores": 1, "kind": _W["kind"],
            "sample": f"{n_done} of {scene.n_envs} environments, one evaluation with derivatives each, single core "
                      f"({t_total:.1f} s CPU); GPU fm/Km/Dm checked on 16 of them: max rel err {worst:.1e}",
            "env_evals_per_sec": n_done / t_total}


def ncu_traffic(key):
    """dram bytes per launch measured by ncu --set full (profiles/ncu_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]["dram_bytes_per_launch"]
    except Exception:
        return None


def flops_per_contact():
    """(flops per active contact of k_force<1,0,0>, source): ncu-measured executed count when profiles/ has one."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "flop_census.json")))["k_force<1,0,0>"]
        return float(d["flops_per_active_contact"]), "ncu executed DFMA*2 + DADD + DMUL per active contact (profiles/flop_census.json)"
    except Exception:
        return float(FLOPS_PER_CONTACT_NF_CENSUS), "hand census of contact_eval_nf + reduction (DESIGN.md §4)"


def c1_leg(ctx, args, torch, stream):
    """env-steps/s of the batched stepper on the C1 configuration (two free parts, ground contact with
    friction and damping + SDF contact, 12 reduced dofs): the general 12-dof kernels and k_ground."""
    from assemble_them_all_b200 import scenes
    n_envs = min(args.envs, 1024)
    rs = scenes.c1_freefall_rollouts(n_envs=n_envs)
    sim, *keep = rs.upload(ctx)
    times = []
    for rep in range(3):
        sim.reset(rs.q0, rs.qdot0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        sim.forward(rs.frame_skip)
        b.record(stream)
        torch.cuda.synchronize()
        if rep:
            times.append(a.elapsed_time(b))
    st = sim.stats()
    out = {"value": n_envs * rs.frame_skip / (1e-3 * float(np.mean(times))), "unit": "env-steps/s", "workload": rs.label,
           "ms_per_forward": float(np.mean(times)), "rounds_per_forward": st["rounds"] / 3.0,
           "converged": int(sim.is_converged().sum()),
           "ground_touched": int(sim.get_ground_history().any(axis=1).sum()),
           "sdf_touched": int(sim.get_contact_history().any(axis=1).sum()),
           "contact_evals_per_env_step": (st["deriv_env_evals"] + st["value_env_evals"]) / max(1, st["env_steps"]),
           "cpu_reference_note": "full-size C1 on the reference: 88 env-steps/s on one core (profiles/r1h_configs_c1_c3.log)"}
    sim.close()
    return out


def rot_leg(ctx, args, torch, stream, n_vertices=10000):
    """env-steps/s of the stepper on the rotational planner's rollouts (free3d-exp joint, 12 actions:
    run_joint_plan.py:585-598): exponential-map kinematics and their derivatives. n_vertices = 10000 is the C2
    geometry, 100000 is C3 (BASELINE.json configs[2]: ~100k contact points per body)."""
    from assemble_them_all_b200 import scenes
    n_envs = min(args.envs, 1024)
    rs = scenes.c2_planner_rollouts(n_envs=n_envs, n_vertices=n_vertices, joint_type="free3d-exp")
    sim, *keep = rs.upload(ctx)
    times = []
    for rep in range(3):
        sim.reset(rs.q0, rs.qdot0)
        for b in range(len(rs.bodies)):
            sim.set_body_external_force(b, rs.f_ext[:, b])
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s0 = sim.stats()
        a.record(stream)
        sim.forward(rs.frame_skip)
        b2.record(stream)
        torch.cuda.synchronize()
        s1 = sim.stats()
        if rep:
            times.append(a.elapsed_time(b2))
    evals = (s1["deriv_env_evals"] + s1["value_env_evals"] - s0["deriv_env_evals"] - s0["value_env_evals"])
    out = {"value": n_envs * rs.frame_skip / (1e-3 * float(np.mean(times))), "unit": "env-steps/s", "workload": rs.label,
           "ms_per_forward": float(np.mean(times)), "rounds_per_forward": s1["rounds"] - s0["rounds"],
           "queries_per_sec_while_stepping": evals * (rs.queries_per_eval // rs.n_envs) / (1e-3 * float(np.mean(times))),
           "converged": int(sim.is_converged().sum()), "touched": int(sim.get_contact_history().any(axis=1).sum())}
    sim.close()
    for o in keep:
        for x in (o if isinstance(o, (list, tuple)) else [o]):
            if hasattr(x, "close"):
                x.close()
    return out


def strong_scaling_leg(ctx, args, rank, world, torch, dist, dev, stream):
    """STRONG scaling of one BFS stage of C2: the FIXED batch of 1024 rollouts x frame_skip 100 is block-partitioned
    over the ranks (1024 / N rollouts per GPU); the timed region covers forward(frame_skip) AND the planner's
    per-stage exchange (one all_gather of the outcome records), max over ranks. What limits it is named in
    DESIGN.md 5: the round loop is latency bound once a GPU holds only 128 rollouts."""
    from assemble_them_all_b200 import scenes, sharding
    import copy
    n_total = 1024
    full = rollout_workload(n_total)
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    rs = copy.copy(full)
    rs.q0, rs.qdot0, rs.f_ext = full.q0[lo:hi], full.qdot0[lo:hi], full.f_ext[lo:hi]
    sim, batch, pairs, sdfs, pts = rs.upload(ctx)
    n = rs.frame_skip

    def prepare():
        sim.reset(rs.q0, rs.qdot0)
        for b in range(len(rs.bodies)):
            sim.set_body_external_force(b, rs.f_ext[:, b])

    prepare()
    sim.forward(n)
    times = []
    for _ in range(3):
        prepare()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        sim.forward(n)
        q, _ = sim.get_state()
        rec = sharding.make_records(np.arange(lo, hi), sim.is_converged().astype(np.float64), np.full(hi - lo, n),
                                    np.arange(lo, hi) % 6, sim.get_contact_history().any(axis=1).astype(np.float64), q)
        table = sharding.gather_records(rec, n_total, dist if world > 1 else None, dev)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    sim.close()
    return {"value": n_total * n / sec, "unit": "env-steps/s", "scaling": "strong", "rollouts_total": n_total,
            "rollouts_per_gpu": hi - lo, "ms_per_stage": 1e3 * sec, "converged": int(table[:, 1].sum()),
            "timing": "wall clock around forward(frame_skip) + state read-back + all_gather of the outcome records, max over ranks",
            "workload": full.label}


def c4_leg(args, torch, local, cpu, rank=0, world=1, dist=None, dev=None):
    """env-steps/s on C4 (BASELINE.json configs[3]): every candidate removal of the 20-part assembly
    (each part x 6 axis directions = 120 rollouts x frame_skip 100; run_multi_plan.py:84-126). A candidate
    has its own joint layout and 38 contact forces, so it is its own batched simulation; the 20 simulations
    are driven concurrently by a few host threads, each with its own library context (stream) and one device
    copy of the 20 grids / point sets. Timed by wall clock between two device synchronisations."""
    import threading
    from assemble_them_all_b200 import contact, scenes
    bodies = scenes.c4_bodies()
    n_workers = max(1, min(len(bodies), int(os.environ.get("SDFC_C4_WORKERS", "10"))))
    ctxs = [contact.Context(local) for _ in range(n_workers)]
    shared = [dict() for _ in range(n_workers)]
    sims = []
    mine = [m for m in range(len(bodies)) if m % world == rank]   # candidates are dealt to the ranks (run_multi_plan.py:84-126)
    n_workers = max(1, min(n_workers, len(mine)))
    for i, m in enumerate(mine):
        w = i % n_workers
        rs = scenes.c4_candidate_rollouts(bodies, m)
        sim, *keep = rs.upload(ctxs[w], shared[w])
        sims.append((w, rs, sim, keep))

    def prepare():
        for _, rs, sim, _ in sims:
            sim.reset(rs.q0, rs.qdot0)
            sim.set_body_external_force(0, rs.f_ext[:, 0])

    def drive(w):
        for ww, rs, sim, _ in sims:
            if ww == w:
                sim.forward(rs.frame_skip)

    times = []
    for rep in range(3):
        prepare()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        th = [threading.Thread(target=drive, args=(w,)) for w in range(n_workers)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        if rep:
            times.append(time.perf_counter() - t0)
    n_roll = sum(rs.n_envs for _, rs, _, _ in sims)
    n = sims[0][1].frame_skip
    sec = float(np.mean(times))
    if world > 1:   # whole job: every candidate of the assembly, time of the slowest rank
        tt = torch.tensor([sec, -float(n_roll)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt[0].item())
        cnt = torch.tensor([float(n_roll)], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        n_roll_total = int(cnt.item())
    else:
        n_roll_total = n_roll
    times = [sec]
    st = [sim.stats() for _, _, sim, _ in sims]
    evals = sum(x["deriv_env_evals"] + x["value_env_evals"] for x in st) / max(1, sum(x["env_steps"] for x in st))
    q_env = sims[-1][1].queries_per_eval // sims[-1][1].n_envs
    # a removal "succeeds" in this leg when the part has moved its own height out along the action
    moved = [float(np.max(np.abs(sim.get_q()), axis=1).max()) for _, _, sim, _ in sims]
    out = {"value": n_roll_total * n / float(np.mean(times)), "unit": "env-steps/s",
           "workload": f"C4 {len(bodies)}-part plate + pegs (~10000 contact pts each), {len(bodies)} candidate removals x 6 "
                       f"axis rollouts x frame_skip {n}, 38 contact forces per rollout",
           "scaling": "strong: the 20 candidates are dealt round-robin to the ranks" if world > 1 else "single GPU",
           "candidates_on_rank0": len(sims),
           "rollouts": n_roll_total, "ms_per_forward": 1e3 * float(np.mean(times)), "host_threads": n_workers,
           "queries_per_env_eval": int(q_env), "contact_evals_per_env_step": evals,
           "converged": int(sum(int(sim.is_converged().sum()) for _, _, sim, _ in sims)),
           "touched": int(sum(int(sim.get_contact_history().any(axis=1).sum()) for _, _, sim, _ in sims)),
           "largest_displacement_per_candidate": [round(x, 3) for x in moved],
           "timing": f"wall clock between device synchronisations ({len(sims)} simulations on {n_workers} streams)"}
    if cpu and have_ref() and world == 1:
        # the reference's Simulation::forward on the same candidate (inputs injected), one rollout, 40 steps (contact from step ~27)
        from oracle import ref_scene
        m, e, k = 7, 0, 40
        w, rs, sim, _ = sims[m]
        sim.reset(rs.q0, rs.qdot0)
        sim.set_body_external_force(0, rs.f_ext[:, 0])
        Q, _ = sim.forward(k, trajectory=True)
        p = ref_scene.make_probe(rs.bodies, rs.kn, rs.kt, rs.mu, rs.damping, inject_frames=True)
        t_ref, q_ref, _ = ref_scene.rollout(p, rs.q0[e], rs.qdot0[e], rs.f_ext[e, 0], k)
        err = float(np.max(np.abs(q_ref - Q[-1, e])) / max(1e-300, np.max(np.abs(q_ref))))
        if err > 1e-9:
            raise SystemExit(f"PARITY FAILURE: C4 candidate {m} rollout {e}: q after {k} steps differs from the reference by {err:.3e}")
        out["cpu_baseline"] = {"value": k / t_ref, "unit": "env-steps/s", "cores": 1, "kind": "reference",
                               "sample": f"candidate part{m}, rollout +x, {k} steps on one core ({t_ref:.1f} s); GPU q after "
                                         f"{k} steps: rel err {err:.1e}"}
    for _, _, sim, _ in sims:
        sim.close()
    for c in ctxs:
        c.close()
    return out


def hbm_regime_leg(ctx, torch, stream, dev, peak, n_points=30_000_000, grid_n=512):
    """The detection kernel where it IS HBM-bound (C5 of BASELINE.json): n_points uniform random points against a
    512^3 grid, one environment. Headline: the library's default for grids of this size - fp32 BRICKS (0.55 GB;
    the tile's voxel brick staged in shared memory with TMA, k_detect_brick) and the k-d point order of
    sdfc_points_create. Beside it: fp32 cell records (4.3 GB, one 32-byte record per query, k_detect_mask) with
    Morton-ordered points (the round-1 configuration) and with points in file order.
    Unit: 24 B point + 32 B eight-corner gather of an fp32-stored grid + 8 B result = 64 B per query
    (96 B, the fp64-storage unit, reported beside it). `traffic`: DRAM bytes of one launch from the committed ncu
    capture (profiles/ncu_traffic.json). CUDA events on the launching stream, L2 flushed between launches."""
    from assemble_them_all_b200 import contact, scenes, synth
    spec, pts = scenes.sweep_scene(n_points, grid_n)
    R = synth.axis_angle_rotation([0.3, -0.2, 0.9], 0.01)
    E = np.stack([synth.make_E(R, [0.01, -0.02, 0.015]), np.eye(4)])[None]
    dE = torch.from_numpy(np.ascontiguousarray(np.swapaxes(E, -1, -2))).to(dev)
    dna = torch.zeros((1, 1), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    runs = {}
    old_layout = os.environ.get("SDFC_GRID_LAYOUT")
    for layout, orders in (("bricks", ("kd",)), ("records", ("morton", "none"))):
        os.environ["SDFC_GRID_LAYOUT"] = layout
        sdf = contact.SdfGrid.from_spec(ctx, spec)
        for order in orders:
            cp = contact.ContactPoints(ctx, pts, morton_sort={"kd": "kd", "morton": True, "none": False}[order])
            pair = contact.ForceGeneralSDFContact(ctx, cp, sdf, 1e6, 1e3)
            batch = contact.ContactBatch(ctx, 1, 2, [pair], [0], [1])
            for _ in range(3):
                batch.detect_device(dE.data_ptr(), dna.data_ptr())
            ms = []
            for _ in range(7):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                batch.detect_device(dE.data_ptr(), dna.data_ptr())
                b.record(stream)
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            runs[(layout, order)] = {"ms": float(np.median(ms)), "active": float(dna.item()) / n_points,
                                     "tiles": batch.brick_stats() if layout == "bricks" else None}
            batch.close(); pair.close(); cp.close()
        sdf.close()
    if old_layout is None:
        os.environ.pop("SDFC_GRID_LAYOUT", None)
    else:
        os.environ["SDFC_GRID_LAYOUT"] = old_layout
    assert len({round(r["active"] * n_points) for r in runs.values()}) == 1, "layouts / orders disagree on the contact count"

    def entry(key, kernel, traffic_key):
        t = runs[key]["ms"] * 1e-3
        tr = ncu_traffic(traffic_key)
        return {"kernel": kernel, "avg_launch_ms": t * 1e3, "queries_per_sec": n_points / t,
                "achieved": n_points * BYTES_PER_QUERY_FP32 / t / 1e9, "frac": n_points * BYTES_PER_QUERY_FP32 / t / 1e9 / peak,
                "frac_96B_unit": n_points * BYTES_PER_QUERY_FP64 / t / 1e9 / peak,
                "traffic": tr, "traffic_bytes_per_query": (tr / n_points) if tr else None}

    head = entry(("bricks", "kd"), "k_detect_brick", "k_detect_brick@C5_3e7x512_kd")
    out = {"bound": "hbm", "peak": peak, "unit": "GB/s", "unit_bytes_per_query": BYTES_PER_QUERY_FP32,
           "algorithmic_bytes_per_launch": n_points * BYTES_PER_QUERY_FP32,
           "workload": f"C5: {n_points:.0e} uniform points vs {grid_n}^3 grid, 1 environment; fp32 bricks (TMA-staged), k-d point order",
           **head, "brick_tiles": runs[("bricks", "kd")]["tiles"],
           "cell_records_morton": entry(("records", "morton"), "k_detect_mask", "k_detect_mask@C5_3e7x512_morton"),
           "cell_records_file_order": entry(("records", "none"), "k_detect_mask", "k_detect_mask@C5_3e7x512_file"),
           "active_fraction": runs[("bricks", "kd")]["active"]}
    del flush
    torch.cuda.empty_cache()
    return out


def plugin_leg(cpu):
    """The drop-in path a RedMax user gets: the reference's OWN redmax_py (unmodified Simulation / Robot / BDF1 /
    Newton / pybind objects), built with the B200 Force plug-ins (oracle/_ref/b200), driven per substep from
    Python the way the BFS planner drives it (tests/drivers/redmax_rollout.py: 6 actions x 100 x forward(1)) on
    the C2 geometry (fastest of 3 passes: a 0.1 s pass of tiny launches is at the mercy of the GPU's clock ramp),
    next to the stock CPU module on the same inputs. Every Robot evaluation is one device call
    (host/B200RobotBatch.h). Wall clock inside forward(); SDF construction excluded (reported by sdf_build)."""
    import glob
    import re
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not (glob.glob(os.path.join(ref, "redmax_py*.so")) and glob.glob(os.path.join(ref, "b200", "redmax_py*.so"))):
        return {"unavailable": "oracle/_ref (stock + b200 redmax_py) not built"}
    driver = os.path.join(ROOT, "tests", "drivers", "redmax_rollout.py")
    out = {"workload": "C2 geometry (2 x ~10000 contact pts, dx 0.05), translational joint, 6 actions x 100 x forward(1) from Python",
           "unit": "env-steps/s"}
    arms = [("b200", os.path.join(ref, "b200"))] + ([("stock_cpu", ref)] if cpu else [])
    finals = {}
    for name, mod in arms:
        with tempfile.TemporaryDirectory() as td:
            npz = os.path.join(td, "o.npz")
            r = subprocess.run([sys.executable, driver, mod, npz, "--nverts", "10000", "--dx", "0.05", "--steps", "100",
                                *(["--reps", "3"] if name == "b200" else [])],
                               capture_output=True, text=True, timeout=900)
            m = re.search(r"\(([0-9.]+) env-steps/s\)", r.stdout)
            if r.returncode != 0 or not m:
                return {"unavailable": f"{name}: " + (r.stdout + r.stderr)[-300:]}
            out[name] = float(m.group(1))
            finals[name] = np.load(npz)["Q"]
    if "stock_cpu" in finals:
        err = float(np.max(np.abs(finals["b200"] - finals["stock_cpu"])) / np.max(np.abs(finals["stock_cpu"])))
        if err > 1e-9:
            raise SystemExit(f"PARITY FAILURE: b200 redmax_py vs stock redmax_py per-substep q differ by {err:.3e}")
        out["max_rel_err_q_vs_stock"] = err
        out["speedup_vs_stock_1core"] = out["b200"] / out["stock_cpu"]
    out["value"] = out["b200"]
    return out


def sdf_build_leg(ctx, cpu):
    """SDF construction (sdfgen::make_level_set3 drop-in) for the two C2 parts: wall time of the host call
    (upload, ~5 k wavefront launches, download); the CPU restatement timed on the larger part."""
    from assemble_them_all_b200 import contact, synth
    parts = synth.peg_in_hole(n_vertices=10000)
    out = {"parts": []}
    for pid in ("0", "1"):
        V, F = parts[pid]
        V = np.float64(np.float32(V))
        contact.precompute_sdf(ctx, V, F, 0.05)   # warm-up
        t0 = time.perf_counter()
        g = contact.precompute_sdf(ctx, V, F, 0.05)
        out["parts"].append({"cells": [g["ni"], g["nj"], g["nk"]], "triangles": int(len(F)), "gpu_ms": 1e3 * (time.perf_counter() - t0)})
    if cpu:
        from oracle import sdfgen_oracle as SG
        V, F = parts["1"]
        V = np.float64(np.float32(V))
        lay = SG.grid_layout(V, 0.05)
        sp = lay["spacing"].astype(np.float32)
        t0 = time.perf_counter()
        ref = SG.make_level_set3(F, V.astype(np.float32), lay["min_box"], sp[0], sp[1], sp[2], *(int(v) for v in lay["dims"]))
        out["cpu_port_ms_part1"] = 1e3 * (time.perf_counter() - t0)
        out["bit_identical_part1"] = bool(np.array_equal(ref, g["grid"].astype(np.float32)))
    return out


def env_steps_leg(ctx, args, rank, world, local, torch, dist, dev, stream):
    """env-steps/s of the batched stepper on the planner rollouts (sdfc_sim_forward = Simulation::forward(frame_skip)
    for every rollout). `value`: state resident in HBM, nothing returned per step. `e2e`: reset + external force
    from host buffers, forward(frame_skip) and the per-substep q / qdot trajectory (what run_joint_plan.py:553
    reads) copied back to the host, all inside the timed region."""
    rs = rollout_workload(args.envs, seed_offset=rank)
    sim, batch, pairs, sdfs, pts = rs.upload(ctx)
    n = rs.frame_skip

    def prepare():
        sim.reset(rs.q0, rs.qdot0)
        for b in range(len(rs.bodies)):
            sim.set_body_external_force(b, rs.f_ext[:, b])

    prepare()
    sim.forward(n)   # warm-up
    reps = max(2, min(args.steps, 5))
    dev_ms, e2e_s = [], []
    st0 = sim.stats()
    l0 = ctx.launch_count
    for _ in range(reps):
        prepare()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        sim.forward(n)
        b.record(stream)
        torch.cuda.synchronize()
        dev_ms.append(a.elapsed_time(b))
    st1 = sim.stats()
    launches = ctx.launch_count - l0
    prepare()
    sim.forward(n, trajectory=True)   # warm-up of the trajectory path (first call allocates the device trajectory buffers)
    for _ in range(reps):
        t0 = time.perf_counter()
        prepare()
        Q, QD = sim.forward(n, trajectory=True)
        e2e_s.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.mean(dev_ms)) * 1e-3, float(np.mean(e2e_s))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(t[0].item()), float(t[1].item())
    steps_total = world * rs.n_envs * n
    d = {k: (st1[k] - st0[k]) / reps for k in st0}
    evals = (d["deriv_env_evals"] + d["value_env_evals"]) / (rs.n_envs * n)
    out = {
        "value": steps_total / t_dev, "unit": "env-steps/s", "workload": rs.label, "rollouts_per_gpu": rs.n_envs,
        "steps_per_rollout": n, "ms_per_forward": 1e3 * t_dev,
        "e2e": {"value": steps_total / t_e2e, "unit": "env-steps/s", "ms_per_forward": 1e3 * t_e2e,
                "h2d_bytes_per_forward": int(rs.q0.nbytes * 2 + rs.f_ext.nbytes * len(rs.bodies)),
                "d2h_bytes_per_forward": int(Q.nbytes + QD.nbytes)},
        "rounds_per_forward": d["rounds"], "contact_evals_per_env_step": evals,
        "queries_per_sec_while_stepping": world * (d["deriv_env_evals"] + d["value_env_evals"]) *
        (rs.queries_per_eval // rs.n_envs) / t_dev,
        "gpu_launches_per_forward": launches / reps,
        "detection_culling": batch.skin_stats(),
    }
    # the planner's only exchange: one all_gather of per-rollout outcome records (NCCL), in rollout order
    from assemble_them_all_b200 import sharding
    n_total = world * rs.n_envs
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    rec = sharding.make_records(np.arange(lo, hi), sim.is_converged().astype(np.float64), np.full(rs.n_envs, n),
                                np.arange(rs.n_envs) % 6, sim.get_contact_history().any(axis=1).astype(np.float64), Q[-1])
    table = sharding.gather_records(rec, n_total, dist if world > 1 else None, dev)
    out["outcomes"] = {"rollouts": int(len(table)), "converged": int(table[:, 1].sum()), "touched": int(table[:, 4].sum()),
                       "gather": "one all_gather of %d x %d fp64 records" % (n_total, sharding.RECORD_WIDTH)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # reference Simulation::forward on one core over a bounded sample of the same rollouts + parity of final states
        _ref_worker_init(scene_blob(rs), have_ref())
        t_cpu, n_cpu, worst_q, done = 0.0, 0, 0.0, 0
        t_start = time.perf_counter()
        for e in range(0, rs.n_envs, max(1, rs.n_envs // 61)):
            tt, ns, kind, finals = _ref_rollout_envs([e])
            t_cpu += tt
            n_cpu += ns
            done += 1
            _, q_ref, _ = finals[0]
            scale = float(np.max(np.abs(q_ref)))
            worst_q = max(worst_q, float(np.max(np.abs(q_ref - Q[-1, e])) / scale) if scale > 0 else 0.0)
            if time.perf_counter() - t_start > 15.0:
                break
        if worst_q > 1e-9:
            raise SystemExit(f"PARITY FAILURE: stepper final q vs CPU {_W['kind']} relative error {worst_q:.3e} > 1e-9")
        out["cpu_baseline"] = {"value": n_cpu / t_cpu, "unit": "env-steps/s", "cores": 1, "kind": _W["kind"],
                               "sample": f"{done} of {rs.n_envs} rollouts x {n} steps on one core ({t_cpu:.1f} s CPU); "
                                         f"GPU final q after {n} steps checked on all of them: max rel err {worst_q:.1e}"}
    sim.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr, so that
    # nothing a library prints (NCCL's version banner comes from C code) can land in front of the line
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the contact path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its version banner (NCCL_DEBUG >= VERSION, WARN included) to stdout; stdout carries ONE JSON
        # line, so the banner level is dropped and whatever NCCL still logs goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from assemble_them_all_b200 import contact
    scene = workload(args.envs, seed_offset=rank)
    ctx = contact.Context(local)
    # an explicit (non-legacy) stream: the library launches on it and the timing
    # events below are recorded on the same stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    batch, pairs, sdfs, pts = scene.upload(ctx)
    nm = batch.ndof_m
    E_col = np.ascontiguousarray(np.swapaxes(scene.E, -1, -2))
    dE = torch.from_numpy(E_col).to(dev)
    dphi = torch.from_numpy(np.ascontiguousarray(scene.phi)).to(dev)
    dfm = torch.empty((scene.n_envs, nm), dtype=torch.float64, device=dev)
    dKm = torch.empty((scene.n_envs, nm * nm), dtype=torch.float64, device=dev)
    dDm = torch.empty((scene.n_envs, nm * nm), dtype=torch.float64, device=dev)
    dna = torch.empty((scene.n_envs, batch.n_pairs), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        batch.eval_device(dE.data_ptr(), dphi.data_ptr(), True, dfm.data_ptr(), dKm.data_ptr(), dDm.data_ptr(), dna.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        flush.fill_(1)
        step()
    # keep the device under this load for >= 1.5 s so that nvidia-smi (100 ms period)
    # samples the clocks the timed steps will run at
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 1.5:
        for _ in range(20):
            step()
        torch.cuda.synchronize()
    barrier()

    # ---- the timed region: K steps, CUDA events around each (the L2 flush stays outside the pair)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    l0 = ctx.launch_count
    for a, b in evs:
        flush.fill_(1)          # L2 flush between timed iterations (outside the event pair)
        a.record(stream)
        step()
        b.record(stream)
    barrier()
    gpu_launches = ctx.launch_count - l0
    per_step = [a.elapsed_time(b) for a, b in evs]
    ms_step = float(np.mean(per_step))
    # ---- the same K steps once more with per-kernel CUDA events (sdfc_context_enable_timing: event pairs around the
    # detection / force / reduction launches on the context's stream; this pass launches kernel by kernel instead
    # of replaying the captured graph, so its step time is a few launch gaps longer than the headline's)
    ctx.enable_timing(True)
    ctx.read_timing(reset=True)
    for _ in range(args.steps):
        flush.fill_(1)
        step()
    barrier()
    kt = ctx.read_timing(reset=True)
    ctx.enable_timing(False)
    for _ in range(40):  # tail load so the sampler also brackets the timed region
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop()

    t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    q_rank = scene.queries_per_eval
    value = world * q_rank / (ms_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry (sdfc_batch_eval): inputs in
    # page-locked host memory, every step copies poses + twists to the device, runs the
    # contact pass and reads fm / Km / contact counts back (Dm is identically zero for
    # the planner's mu = 0, damping = 0 and is filled on the host without a transfer)
    pin = {"E": contact.PinnedArray((scene.n_envs, batch.n_bodies, 16)), "phi": contact.PinnedArray((scene.n_envs, batch.n_bodies, 6)),
           "fm": contact.PinnedArray((scene.n_envs, nm)), "Km": contact.PinnedArray((scene.n_envs, nm, nm)),
           "Dm": contact.PinnedArray((scene.n_envs, nm, nm)), "na": contact.PinnedArray((scene.n_envs, batch.n_pairs), np.int32)}
    pin["E"].a[:] = E_col.reshape(scene.n_envs, batch.n_bodies, 16)
    pin["phi"].a[:] = scene.phi

    def e2e_step():
        batch.eval_into(pin["E"], pin["phi"], True, pin["fm"], pin["Km"], pin["Dm"], pin["na"])

    for _ in range(3):
        e2e_step()
    barrier()
    e2e_times = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_step()
        e2e_times.append(time.perf_counter() - t0)
    barrier()
    t = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * q_rank / float(t.item())
    h2d = scene.n_envs * batch.n_bodies * (16 + 6) * 8
    d2h = scene.n_envs * (nm + nm * nm) * 8 + scene.n_envs * batch.n_pairs * 4
    # (n_envs, nm, nm) column-major -> row-major views for the checker below
    out = (pin["fm"].a.copy(), np.swapaxes(pin["Km"].a, 1, 2).copy(), np.swapaxes(pin["Dm"].a, 1, 2).copy(), pin["na"].a.copy())
    # the pageable-buffer form of the same call (numpy arrays in, new arrays out) must agree
    chk = batch.eval(scene.E, scene.phi, deriv=True)
    assert all(np.array_equal(a, b) for a, b in zip(out, chk)), "pinned and staged host entries disagree"

    # ---- final gather of outcome records (the planner's only exchange)
    n_active = dna.sum(dim=0).to(torch.int64)
    if world > 1:
        gathered = [torch.zeros_like(n_active) for _ in range(world)]
        dist.all_gather(gathered, n_active)
        n_active = torch.stack(gathered).sum(dim=0)
    contacts_per_step = int(n_active.sum().item())

    env_steps = None if args.skip_env_steps else env_steps_leg(ctx, args, rank, world, local, torch, dist, dev, stream)
    extras = {}
    if not args.skip_env_steps:
        # every rank takes part: the fixed 1024-rollout stage split over the ranks, the 20 C4 candidates dealt to them
        extras["c2_strong_scaling"] = strong_scaling_leg(ctx, args, rank, world, torch, dist, dev, stream)
        extras["c4_candidate_removals"] = c4_leg(args, torch, local, cpu=(world == 1 and not args.no_cpu_baseline),
                                                 rank=rank, world=world, dist=dist, dev=dev)
    if rank == 0 and not args.skip_env_steps:
        extras["c1_freefall"] = c1_leg(ctx, args, torch, stream)
        extras["rotational_rollouts"] = rot_leg(ctx, args, torch, stream)
        extras["c3_rotational_100k"] = rot_leg(ctx, args, torch, stream, n_vertices=100000)
        extras["sdf_build"] = sdf_build_leg(ctx, cpu=(world == 1 and not args.no_cpu_baseline))
        if world == 1:
            extras["plugin_drop_in"] = plugin_leg(cpu=not args.no_cpu_baseline)
        extras["roofline_hbm_regime"] = hbm_regime_leg(ctx, torch, stream, dev, measured_hbm_peak()[0])

    live = ctx.measure_peaks()
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        det_ms = kt["detect"][0] / max(1, kt["detect"][1])
        frc_ms = kt["force"][0] / max(1, kt["force"][1])
        red_ms = (kt["reduce"][0] + kt["scatter"][0]) / max(1, kt["reduce"][1])
        fp32 = all(s_.is_fp32 for s_ in sdfs)
        unit_b = BYTES_PER_QUERY_FP32 if fp32 else BYTES_PER_QUERY_FP64
        fpc, fpc_src = flops_per_contact()
        contacts_rank = contacts_per_step / world
        force_tflops = contacts_rank * fpc / (frc_ms * 1e-3) / 1e12
        det_l2 = q_rank * unit_b / (det_ms * 1e-3) / 1e9
        # k_force<1,0,0> (planner parameters: mu = 0, damping = 0), bounded by the FP64 pipe: achieved = active contacts x
        # flops per contact / average launch time (CUDA events, second pass of the timed steps; plan kernel included);
        # peak = DFMA throughput measured on this device in this process (sdfc_measure_peaks).
        rl_force = {"kernel": "k_force<1,0,0>", "bound": "fp64", "achieved": force_tflops, "peak": live["fp64_gflops"] / 1e3,
                         "unit": "TFLOP/s", "frac": force_tflops / (live["fp64_gflops"] / 1e3), "traffic": ncu_traffic("k_force@C2"),
                         "peak_source": "measured live: sdfc_measure_peaks (8 independent DFMA chains per thread, all SMs)",
                         "algorithmic_flops_per_launch": contacts_rank * fpc, "flops_per_active_contact": fpc,
                         "flops_source": fpc_src, "active_contacts_per_launch": contacts_rank, "avg_launch_ms": frc_ms,
                         "share_of_step": frc_ms / (det_ms + frc_ms + red_ms),
                         "note": "bound 'fp64' is outside the contract's hbm|tensor enum on purpose: this kernel moves ~3 MB of "
                                 "DRAM per launch and has no tensor-core work; it is latency / occupancy bound (DESIGN.md 4)"}
        # Detection on C2 is cache resident (2 point sets + 2 grids shared by all rollouts): its ceiling is L2 -> SM
        # delivery (measured live), not HBM. A fraction above 1 means L1 hits on top of L2.
        rl_detect = {"kernel": "k_detect_mask", "bound": "l2", "achieved": det_l2, "peak": live["l2_read_gbs"],
                                   "unit": "GB/s", "frac": det_l2 / live["l2_read_gbs"], "traffic": ncu_traffic("k_detect_mask@C2"),
                                   "unit_bytes_per_query": unit_b, "frac_96B_unit": q_rank * BYTES_PER_QUERY_FP64 / (det_ms * 1e-3) / 1e9 / live["l2_read_gbs"],
                                   "peak_source": "measured live: sdfc_measure_peaks (every CTA streams a 48 MiB buffer, 128-bit loads)",
                                   "hbm_peak": peak, "hbm_peak_source": peak_src,
                                   "algorithmic_bytes_per_launch": q_rank * unit_b, "avg_launch_ms": det_ms,
                                   "share_of_step": det_ms / (det_ms + frc_ms + red_ms)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": scene.label, "envs_per_gpu": scene.n_envs, "pairs": batch.n_pairs,
                       "queries_per_step_per_gpu": q_rank, "active_contacts_per_step": contacts_per_step,
                       "derivatives": True, "sdf_storage": "fp32-lossless" if all(s.is_fp32 for s in sdfs) else "fp64",
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "parity_mode": "bit-critical arithmetic in non-contracted IEEE intrinsics",
                       "roofline_note": "C2 shares 2 point sets + 2 grids (~65 MB) across all rollouts: detection is L1/L2-served "
                                        "(roofline_detect_c2, against the measured L2 peak), the step's dominant kernel is FP64-bound "
                                        "(roofline); the HBM-bound regime of detection is roofline_hbm_regime (C5)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * float(np.mean(e2e_times))},
            "gpu_launches": gpu_launches,
            # per-kernel rooflines of the step (filled below): "roofline" is a copy of the time-dominant one
            "roofline": None, "roofline_force_c2": None, "roofline_detect_c2": None,
            "peaks_measured_live": live,
            "kernels_ms": {"k_detect_mask": det_ms, "k_force_plan+k_force": frc_ms, "k_reduce_scatter": red_ms,
                           "step": ms_step,
                           "note": "per-kernel figures: CUDA events around each launch in a second pass of the same K steps "
                                   "(kernel-by-kernel launches); 'step' is the headline pass (captured CUDA graph replayed)"},
            "env_evals_per_sec": world * scene.n_envs / (ms_max * 1e-3),
            "env_steps": env_steps,
            **extras,
        }
        line["roofline_force_c2"], line["roofline_detect_c2"] = rl_force, rl_detect
        line["roofline"] = dict(rl_detect if det_ms >= frc_ms else rl_force)
        line["roofline"]["dominant_because"] = "largest share of the step among its kernels (kernels_ms)"
        if world == 1 and not args.no_cpu_baseline:
            fm_h, Km_h, Dm_h, _ = out
            line["cpu_baseline"] = cpu_baseline_leg(scene, (fm_h, Km_h, Dm_h))
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=1024, help="environments per GPU")
    ap.add_argument("--ref-envs-per-core", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-env-steps", action="store_true", help="tuning runs: contact pass only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
