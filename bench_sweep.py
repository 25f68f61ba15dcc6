#!/usr/bin/env python
"""bench_sweep.py — C5 of BASELINE.json: contact-query throughput sweep, points x SDF grid size.

    python bench_sweep.py [--points 1e6,1e7,1e8] [--grids 64,128,256,512] [--reps 5] [--out profiles/x.jsonl]

This is the HBM-bound regime of the query kernels (bench.py's C2 workload is cache
resident by construction). Points are uniform in the grid box (no early-outs),
seeded; the grid is the analytic sphere SDF rounded to fp32 (fresh-grid storage).
Two entries are timed with CUDA events, inputs resident in HBM, a 256 MiB L2 flush
between repetitions:

  distance   sdfc_sdf_distance_device  = SDFMesh.distance (Mesh.cpp:143-198): one
             trilinear sample per point, fp64 result per point
  detect     sdfc_batch_detect_device  = collision_detection_general_SDF: two rigid
             transforms + sample + (d < 0) -> 1 bit per point

Bytes per query: ALGORITHMIC 96 B (24 B point + 64 B fp64 corner
gather + 8 B result) and COMPULSORY DRAM bytes of the fp32 cell-record layout when
the grid exceeds L2 (24 B point + 32 B record sector + 8 B result = 64 B for
`distance`; 24 + 32 + 1/8 B for `detect`). One JSON line per (entry, N, grid).
Multi-GPU (torchrun): points are sharded, the grid is replicated, no collective.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="1e6,1e7,1e8")
    ap.add_argument("--grids", default="64,128,256,512")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--active-frac", type=float, default=0.05)
    ap.add_argument("--out", default="")
    ap.add_argument("--order", choices=["none", "morton", "kd"], default="kd",
                    help="detect: point order of the one-time sort sdfc_points_create does (none = file order)")
    ap.add_argument("--no-morton", action="store_true", help="same as --order none")
    ap.add_argument("--layout", choices=["auto", "records", "bricks", "both"], default="auto",
                    help="device storage of fp32-lossless grids (auto: cell records up to ~96 MB of records, bricks beyond)")
    ap.add_argument("--entries", default="distance,detect")
    args = ap.parse_args()
    if args.no_morton:
        args.order = "none"
    if args.layout != "auto":
        os.environ["SDFC_GRID_LAYOUT"] = args.layout
    import torch
    import torch.distributed as dist
    from assemble_them_all_b200 import contact, scenes, synth
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    ctx = contact.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    lines = []
    for gn in [int(float(x)) for x in args.grids.split(",")]:
        spec = synth.sphere_grid(gn)
        rng = np.random.default_rng(synth.SWEEP_SEED + rank)
        probe = rng.uniform(spec.min_box * 0.999, spec.max_box * 0.999, size=(200000, 3))
        spec.col_th = float(-np.quantile(np.linalg.norm(probe, axis=1) - 3.5, args.active_frac))
        sdf = contact.SdfGrid.from_spec(ctx, spec)
        layout = sdf.layout
        grid_bytes = gn ** 3 * {"records": 32, "bricks": 4, "bricks+records": 36, "fp64": 8}[layout]
        corner_bytes = 32 if sdf.is_fp32 else 64   # the gather is 32 B when the grid is stored fp32
        for npts in [int(float(x)) for x in args.points.split(",")]:
            n_local = npts // world
            pts = rng.uniform(spec.min_box * 0.999, spec.max_box * 0.999, size=(n_local, 3))
            soa = torch.from_numpy(np.ascontiguousarray(pts.T)).to(dev)   # x[n] y[n] z[n]
            dout = torch.empty(n_local, dtype=torch.float64, device=dev)
            # ---- distance
            def run_distance():
                sdf.distance_device(n_local, soa[0].data_ptr(), soa[1].data_ptr(), soa[2].data_ptr(), dout.data_ptr())
            # ---- detect (1 environment, identity-ish poses with a small seeded rotation)
            cp = contact.ContactPoints(ctx, pts, morton_sort={"none": False, "morton": True, "kd": "kd"}[args.order])
            pair = contact.ForceGeneralSDFContact(ctx, cp, sdf, 1e6, 1e3)
            batch = contact.ContactBatch(ctx, 1, 2, [pair], [0], [1])
            R = synth.axis_angle_rotation([0.3, -0.2, 0.9], 0.01)
            E = np.stack([synth.make_E(R, [0.01, -0.02, 0.015]), np.eye(4)])[None]
            dE = torch.from_numpy(np.ascontiguousarray(np.swapaxes(E, -1, -2))).to(dev)
            dna = torch.zeros((1, 1), dtype=torch.int32, device=dev)

            def run_detect():
                batch.detect_device(dE.data_ptr(), dna.data_ptr())

            for name, fn, comp in (("distance", run_distance, 24 + 32 + 8), ("detect", run_detect, 24 + 32 + 0.125)):
                if name not in args.entries.split(","):
                    continue
                for _ in range(3):
                    fn()
                ms = []
                for _ in range(args.reps):
                    flush.fill_(1)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    fn()
                    b.record(stream)
                    torch.cuda.synchronize()
                    ms.append(a.elapsed_time(b))
                t = torch.tensor([float(np.median(ms))], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item()) * 1e-3
                qps = n_local * world / sec
                line = {"entry": name, "points": npts, "grid": gn, "n_gpus": world, "ms": sec * 1e3, "queries_per_sec": qps,
                        # SURVEY 8(d) unit: 24 B point + 8-corner gather (32 B fp32 storage / 64 B fp64) + 8 B result
                        "unit_bytes": 24 + corner_bytes + 8,
                        "algorithmic_GBs": qps * (24 + corner_bytes + 8) / 1e9,
                        "frac_of_hbm_peak": qps * (24 + corner_bytes + 8) / 1e9 / (peak * world),
                        "frac_of_hbm_peak_96B_unit": qps * 96 / 1e9 / (peak * world),
                        "hbm_peak_GBs_per_gpu": peak, "grid_storage": layout,
                        "grid_bytes": grid_bytes, "fits_l2": grid_bytes < 120e6,
                        "point_order": ("as given" if name == "distance" else {"none": "file order", "morton": "Morton-sorted at upload",
                                                                               "kd": "k-d sorted at upload"}[args.order]),
                        "active_fraction": float(dna.item()) / n_local if name == "detect" else None,
                        "brick_tiles": batch.brick_stats() if name == "detect" and layout.startswith("bricks") else None,
                        "brick_smem": os.environ.get("SDFC_BRICK_SMEM")}
                if rank == 0:
                    print(json.dumps(line), flush=True)
                    lines.append(line)
            batch.close(); pair.close(); cp.close()
            del soa, dout, pts
            torch.cuda.empty_cache()
        sdf.close()
    if rank == 0 and args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
