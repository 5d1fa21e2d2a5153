"""Host-buffer (e2e) verify throughput for both pipeline schedules and several chunk sizes (KB_VERIFY_PIPE, KB_VERIFY_CHUNK).
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
# (pipeline schedule, chunk size): KB_VERIFY_PIPE = 1 one compute stream + one copy stream, 0 two independent lanes;
# chunk sizes as powers of two and as multiples of 56832 (a wave of the main kernel when it ran 3 blocks per SM; 75776 at 4)
cases = [(None, None)] + [(p, c) for p in (0, 1) for c in (227328, 454656, 909312, 1 << 20)] + [(None, None)]
for pipe, chunk in cases:
    for key, val in (("KB_VERIFY_PIPE", pipe), ("KB_VERIFY_CHUNK", chunk)):
        if val is None:
            os.environ.pop(key, None)   # the library's defaults
        else:
            os.environ[key] = str(val)
    ctx = kb.Context(0)
    for _ in range(2):
        ctx.verify_batch(h_pk, h_msg, h_off, h_sig, out=h_out)
    best = 1e9
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        t1 = time.perf_counter()
        h_out[:] = 0xff
        ctx.verify_batch(h_pk, h_msg, h_off, h_sig, out=h_out)
        best = min(best, time.perf_counter() - t1)
        assert (h_out == expect).all()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"pipe": pipe, "chunk": chunk, "e2e_sigs_per_s": n / dt, "ms": dt * 1e3, "best_ms": best * 1e3}), flush=True)
    ctx.close()
