"""The N>1 path on CPU: world_size 2 over gloo.  Work is sharded by utterance (u mod G), no data-path collective;
only the 53-dim rows are gathered on the host.  The oracle stands in for the per-rank compute here."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from webspeechanalyzer_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_utt, q):
    import torch.distributed as dist

    from oracle import oracle
    from webspeechanalyzer_b200 import FaConfig, synth_speech
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = FaConfig.default(output_level=13)
    mine = shard.shard_indices(n_utt, rank, world)
    rows = []
    for u in mine:
        _, an = oracle.analyze_pcm(cfg, synth_speech(3 * 16000, 16000, 5, int(u)), 16000)
        rows.append(an.features)
    keys, r = shard.pack_rows(mine, rows)
    out = shard.gather_rows(keys, r, dst=0)
    if rank == 0:
        q.put((out[0], out[1]))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition():
    for n in (0, 1, 7, 100):
        for g in (1, 2, 4, 8):
            parts = [shard.shard_indices(n, r, g) for r in range(g)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gather_equals_single_process():
    from oracle import oracle
    from webspeechanalyzer_b200 import FaConfig, synth_speech
    n_utt = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_utt, q)) for r in range(2)]
    for p in procs:
        p.start()
    keys, rows = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = FaConfig.default(output_level=13)
    ref = [oracle.analyze_pcm(cfg, synth_speech(3 * 16000, 16000, 5, u), 16000)[1].features for u in range(n_utt)]
    k2, r2 = shard.pack_rows(range(n_utt), ref)
    assert np.array_equal(keys, k2) and np.array_equal(rows, r2, equal_nan=True)
