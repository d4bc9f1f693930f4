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


def _worker(rank, world, port, n_utt, q, level=13):
    import torch.distributed as dist

    from oracle import oracle
    from webspeechanalyzer_b200 import FaConfig, synth_speech
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host_group = dist.new_group(backend="gloo")     # what an NCCL job uses for the host-side gather
    cfg = FaConfig.default(output_level=level)
    mine = shard.shard_indices(n_utt, rank, world)
    rows = []
    for u in mine:
        _, an = oracle.analyze_pcm(cfg, synth_speech(3 * 16000, 16000, 5, int(u)), 16000)
        rows.append(an.utterance if level == 11 else an.features)
    width = 264 if level == 11 else 53
    # the dense layout of Engine.result(None): rows of the shard's utterances back to back + rows per utterance
    dense = np.concatenate(rows) if rows else np.zeros((0, width))
    keys = shard.keys_from_counts(mine, [len(x) for x in rows])
    k2, r2 = shard.pack_rows(mine, rows, width)
    assert np.array_equal(keys, k2) and np.array_equal(dense, r2, equal_nan=True)
    out = shard.gather_rows(keys, dense, dst=0, group=host_group)
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


import pytest  # noqa: E402


@pytest.mark.parametrize("level", [13, 11])
def test_two_rank_gather_equals_single_process(level):
    from oracle import oracle
    from webspeechanalyzer_b200 import FaConfig, synth_speech
    n_utt = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_utt, q, level)) for r in range(2)]
    for p in procs:
        p.start()
    keys, rows = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = FaConfig.default(output_level=level)
    ans = [oracle.analyze_pcm(cfg, synth_speech(3 * 16000, 16000, 5, u), 16000)[1] for u in range(n_utt)]
    ref = [a.utterance if level == 11 else a.features for a in ans]
    k2, r2 = shard.pack_rows(range(n_utt), ref, 264 if level == 11 else 53)
    assert rows.shape[1] == (264 if level == 11 else 53) and len(rows) > 0
    assert np.array_equal(keys, k2) and np.array_equal(rows, r2, equal_nan=True)
