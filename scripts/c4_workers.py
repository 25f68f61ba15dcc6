"""C4 leg of bench.py alone, for several host-thread counts (SDFC_C4_WORKERS). GPU box only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    for w in (sys.argv[1:] or ["4", "10", "20"]):
        os.environ["SDFC_C4_WORKERS"] = w
        out = bench.c4_leg(None, torch, 0, cpu=False)
        print(w, json.dumps({k: out[k] for k in ("value", "ms_per_forward", "host_threads", "converged", "touched")}), flush=True)
