"""Column sharding: columns are independent, so a grid is split into
contiguous column blocks — one per GPU (or per rank) — with no collective on the data path."""
import threading


def column_blocks(ncol, nshards, align=128):
    """Split ``[0, ncol)`` into ``nshards`` contiguous blocks ``(start, stop)``.

    Block sizes are multiples of ``align`` (the kernels' CTA width) except the last; blocks
    differ by at most ``align`` columns; empty blocks are returned as ``(ncol, ncol)`` when
    there are fewer aligned units than shards.
    """
    if nshards < 1:
        raise ValueError('nshards must be >= 1')
    if ncol < 0:
        raise ValueError('ncol must be >= 0')
    units = (ncol + align - 1) // align
    base, extra = divmod(units, nshards)
    blocks, u0 = [], 0
    for r in range(nshards):
        u1 = u0 + base + (1 if r < extra else 0)
        blocks.append((min(u0 * align, ncol), min(u1 * align, ncol)))
        u0 = u1
    return blocks


def rank_block(ncol, rank, world_size, align=128):
    """The block owned by ``rank`` of ``world_size`` (torchrun: one process per GPU)."""
    return column_blocks(ncol, world_size, align)[rank]


def run_on_devices(fn, blocks, devices):
    """Run ``fn(block, device)`` for every (block, device) pair on its own host thread
    (ctypes releases the GIL for the duration of the C call).  Exceptions are re-raised."""
    errs = []

    def work(blk, dev):
        try:
            fn(blk, dev)
        except BaseException as e:  # noqa: BLE001 - re-raised below
            errs.append(e)

    ths = [threading.Thread(target=work, args=(b, d)) for b, d in zip(blocks, devices) if b[1] > b[0]]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errs:
        raise errs[0]
