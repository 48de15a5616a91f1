"""World-size-2 run of the segment-parallel host logic on CPU (gloo).  Each rank owns its share of
the segment windows, runs them (through the oracle here — there is no GPU in this container), and
the gathered per-segment results must equal a single-process pass.  The data path has no collective;
gloo is only the test's transport for comparing results, as NCCL is only bench.py's timing barrier."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fun_asr_gguf_b200 import segments, weights as Wm
from oracle import oracle as O
from tests import signals

SR = 16000


def _run_segments(audio, wins, owned, w, c):
    out = {}
    for batch_idx in segments.pack_batches([wins[i][1] - wins[i][0] for i in owned], 1):
        batch, lens = segments.pad_batch(audio, [wins[i] for i in owned], batch_idx)
        enc, _ = O.encode_batch(torch.from_numpy(batch), lens, w, c)
        ids = O.ctc_ids_batch(enc, w)
        for r, j in enumerate(batch_idx):
            out[owned[j]] = (ids[r].numpy(), lens[r])
    return out


def _worker(rank, world, port, audio, wins, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    w, c = Wm.random_weights(0), Wm.front_end_constants(64)
    mine = _run_segments(audio, wins, segments.shard(len(wins), world, rank), w, c)
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: (v[0].tolist(), v[1]) for k, v in mine.items()})
    dist.barrier()
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def test_two_ranks_cover_all_segments_and_match_one_rank():
    audio = signals.structured(5 * SR, 17).numpy()
    wins = segments.segment_windows(len(audio), segment_s=1.5, overlap_s=0.25)
    assert len(wins) >= 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, audio, wins, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for part in gathered:
        assert not (set(part) & set(merged))          # no segment done twice
        merged.update(part)
    assert sorted(merged) == list(range(len(wins)))   # none dropped
    w, c = Wm.random_weights(0), Wm.front_end_constants(64)
    single = _run_segments(audio, wins, list(range(len(wins))), w, c)
    for i in range(len(wins)):
        # batches of one segment: the physical length (which the unmasked CTC head's ids depend on)
        # is the segment's own, so who owns a segment cannot change its ids
        assert merged[i][1] == single[i][1] == wins[i][1] - wins[i][0]
        assert merged[i][0] == single[i][0].tolist()
