"""Host-buffer (e2e) verify throughput for several pipeline chunk sizes (KB_VERIFY_CHUNK_LOG2).
Usage: python tools/e2e_sweep.py [log2n] — prints one JSON line per chunk size."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

kb = importlib.import_module("kyber-rs_b200")
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log2n
ctx0 = kb.Context(0)
pk, msg, off, sig, expect = bench.make_batch(ctx0, n, 0)
hp = [torch.from_numpy(x).pin_memory() for x in (pk, msg, off.view(np.int64), sig)]
h_pk, h_msg, h_off, h_sig = [x.numpy() for x in hp]
h_off = h_off.view(np.uint64)
h_out = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
for lg in (15, 16, 17, 18, 19):
    os.environ["KB_VERIFY_CHUNK_LOG2"] = str(lg)
    ctx = kb.Context(0)
    for _ in range(2):
        ctx.verify_batch(h_pk, h_msg, h_off, h_sig, out=h_out)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ctx.verify_batch(h_pk, h_msg, h_off, h_sig, out=h_out)
    dt = (time.perf_counter() - t0) / reps
    assert (h_out == expect).all()
    print(json.dumps({"chunk_log2": lg, "e2e_sigs_per_s": n / dt, "ms": dt * 1e3}))
    ctx.close()
