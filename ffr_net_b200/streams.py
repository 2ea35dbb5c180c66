"""Chunked concurrent execution of the frozen (eval-mode) forward passes.

Every kernel of the forward path is a persistent grid of one CTA per SM with a static tile schedule, so a launch whose
tile count is not a multiple of 148 leaves most SMs idle during its last wave (900 tiles for the dominant 256->256@14x14
convolution at batch 512 = 6.08 waves), and the HBM-bound passes between the convolutions (SE gate + residual, stem,
exports) leave the tensor pipe idle. Eval-mode forwards are independent per image, so the batch is cut into a few
chunks, each chunk's kernel sequence goes to its own CUDA stream, and the hardware block scheduler fills the SMs one
chunk's kernel frees with the other chunk's next kernel. Chunk 0 stays on the caller's stream; the side streams fork
from it and join back through events, which also makes the pattern capturable in a CUDA graph (trainer.capture_step).

FFR_STREAMS sets the number of concurrent chunks (a chunk is never smaller than FFR_MIN_CHUNK images). The default
is 1: measured on a B200 at batch 512 (tools/streams_sweep.py, profiles/r01_streams_sweep_512.json) 2 chunks give the
same throughput as 1 (8.51 vs 8.45 ms backbone, 11.44 vs 11.09 ms with RecNet) and 3-4 chunks are 5-8 % slower — the
half-batch kernels lose to their own tail waves (and extra launches) what the overlap wins. Kept as an opt-in.
"""
import os

import torch

_side = {}


def num_streams():
    return max(1, int(os.environ.get("FFR_STREAMS", "1")))


def min_chunk():
    return max(1, int(os.environ.get("FFR_MIN_CHUNK", "64")))


def chunk_bounds(n, k=None):
    """[(lo, hi)] covering range(n): at most `k` near-equal chunks of at least min_chunk() images."""
    k = num_streams() if k is None else k
    k = max(1, min(k, n // min_chunk()))
    base, rem = divmod(n, k)
    out, lo = [], 0
    for i in range(k):
        hi = lo + base + (1 if i < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _side_stream(i, device):
    key = (i, str(device))
    s = _side.get(key)
    if s is None:
        s = _side[key] = torch.cuda.Stream(device=device)
    return s


def fork_join(bounds, device, fn):
    """fn(i, lo, hi) for every chunk: chunk 0 on the current stream, chunk i >= 1 on side stream i. Inputs must have
    been produced on (or be visible to) the current stream; on return all chunks are ordered before later work on it."""
    if len(bounds) == 1:
        fn(0, bounds[0][0], bounds[0][1])
        return
    main = torch.cuda.current_stream(device)
    ready = torch.cuda.Event()
    ready.record(main)
    done = []
    for i in range(1, len(bounds)):
        s = _side_stream(i, device)
        s.wait_event(ready)
        with torch.cuda.stream(s):
            fn(i, bounds[i][0], bounds[i][1])
            e = torch.cuda.Event()
            e.record(s)
        done.append(e)
    fn(0, bounds[0][0], bounds[0][1])
    for e in done:
        main.wait_event(e)
