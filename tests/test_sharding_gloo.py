"""N > 1 host logic on CPU: two gloo ranks each own a contiguous column block
(xcape_b200.sharding.rank_block — the partition bench.py / the multi-GPU path use), compute it
independently (here with the oracle standing in for the device), and the concatenation of the
blocks equals the single-process result.  No data-path collective exists; the
only collectives are the bench's barrier and max-over-ranks timing, exercised here too."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import oracle
    from xcape_b200.sharding import rank_block
    from xcape_b200.synthetic import make_soundings

    dist.init_process_group('gloo', rank=rank, world_size=world)
    ncol = 1000
    c0, c1 = rank_block(ncol, rank, world)
    d = make_soundings('C1', cols=(c0, c1))          # each rank regenerates only its own shard
    out = oracle.calc_cape_ref(d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'], source='most-unstable',
                               pinc=500., vertical_lev='sigma', tmode=oracle.SPEC)
    np.savez(os.path.join(tmp, f'rank{rank}.npz'), c0=c0, c1=c1, cape=out[0], cin=out[1], mu=out[2], z=out[3])
    # bench.py's timing reduction: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    n = torch.tensor([c1 - c0], dtype=torch.int64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    assert n.item() == ncol
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_column_sharding(tmp_path, oracle_mod):
    import torch.multiprocessing as mp
    from xcape_b200.synthetic import make_soundings
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    d = make_soundings('C1')
    full = oracle_mod.calc_cape_ref(d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'], source='most-unstable',
                                    pinc=500., vertical_lev='sigma', tmode=oracle_mod.SPEC)
    parts = [np.load(tmp_path / f'rank{r}.npz') for r in range(world)]
    assert parts[0]['c0'] == 0 and parts[0]['c1'] == parts[1]['c0'] and parts[1]['c1'] == 1000
    assert int(parts[0]['c1']) % 128 == 0
    for key, ref in zip(('cape', 'cin', 'mu', 'z'), full):
        assert np.array_equal(np.concatenate([q[key] for q in parts]), ref), key
