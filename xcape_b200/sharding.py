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


def bind_host_to_gpu(device):
    """Restrict the calling process to the CPUs of the NUMA node ``device`` hangs off, so that the
    pinned staging buffers it allocates afterwards (first touch) and its copy threads sit next to
    that GPU's PCIe root — with one process per GPU on a two-socket host the H2D streams otherwise
    cross the socket interconnect.  Returns the node number, or ``None`` when the topology is not
    visible (containers without sysfs NUMA data, single-node hosts) and nothing was changed."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            raise OSError('no NUMA node in sysfs')
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = cpus & allowed
        if not cpus or cpus == allowed:
            return None if not cpus else node
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, AttributeError, ValueError, RuntimeError, ImportError):
        pass
    # containers often hide the NUMA node in sysfs; the driver still knows each GPU's CPU affinity
    try:
        cpus = gpu_cpu_affinity(device)
        allowed = os.sched_getaffinity(0)
        if cpus is None or not (cpus & allowed) or (cpus & allowed) == allowed:
            return None
        os.sched_setaffinity(0, cpus & allowed)
        return 'cpus ' + format_cpulist(cpus & allowed)
    except (OSError, AttributeError, ValueError):
        return None


def gpu_cpu_affinity(device, topo_text=None):
    """CPU set next to GPU ``device`` from the "CPU Affinity" column of ``nvidia-smi topo -m`` (None if unavailable)."""
    import re
    import subprocess
    if topo_text is None:
        try:
            topo_text = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True, timeout=20).stdout
        except (OSError, subprocess.SubprocessError):
            return None
    topo_text = re.sub(r'\x1b\[[0-9;]*m', '', topo_text)
    header = None
    for line in topo_text.splitlines():
        cells = [c.strip() for c in line.split('\t')]
        if header is None and 'CPU Affinity' in cells:
            header = cells
            continue
        if header is not None and cells and cells[0] == f'GPU{device}':
            # data rows carry the row label in column 0; the header row starts with an empty cell
            idx = header.index('CPU Affinity')
            if idx < len(cells) and re.fullmatch(r'[0-9,\- ]+', cells[idx] or 'x'):
                return parse_cpulist(cells[idx])
    return None


def format_cpulist(cpus):
    cpus = sorted(cpus)
    out, i = [], 0
    while i < len(cpus):
        j = i
        while j + 1 < len(cpus) and cpus[j + 1] == cpus[j] + 1:
            j += 1
        out.append(str(cpus[i]) if i == j else f'{cpus[i]}-{cpus[j]}')
        i = j + 1
    return ','.join(out)


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (sysfs cpulist format)."""
    out = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        out.update(range(int(lo), int(hi or lo) + 1))
    return out
