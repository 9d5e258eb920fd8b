"""Page-locked host buffers for the host-facing step (BatchedEvergladesEnv.step_host), placed next to the GPU.

The end-to-end path is bound by the host link: every step moves the action rows host->device and the step's results
device->host.  With several GPUs per box the results of all ranks land in host memory at once, so WHERE the pinned
pages live matters: ``pinned_empty`` maps anonymous memory, binds it to the NUMA node the GPU's PCIe root hangs off
(``mbind``; the node is read from sysfs), touches it and registers it with CUDA (``cudaHostRegister``).  When any of
that is not permitted (single-node box, cpuset without that node, seccomp) it falls back to torch's pinned allocator.
``EVG_HOST_NUMA=0`` forces the fallback.  Plumbing only: nothing here touches game state.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import os

_keep = []          # (address, bytes) of every mapping handed out: they live as long as the process
_last = {"method": None, "numa_node": None}

_SYS_MBIND = 237     # x86_64
_MPOL_BIND, _MPOL_PREFERRED = 2, 1
_MADV_HUGEPAGE = 14
_PROT_RW, _MAP_PRIVATE_ANON = 0x3, 0x22


def last_placement() -> dict:
    """How the most recent pinned_empty() placed its pages: {'method': 'mbind+register' | 'torch', 'numa_node': int | None}."""
    return dict(_last)


def gpu_numa_node(device_index: int):
    """NUMA node of the GPU's PCI function (sysfs), or None when the box does not say (single node, VM without topology)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def _numa_nodes_online() -> int:
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except Exception:
        return 1


def _bound_mapping(nbytes: int, node: int):
    libc = C.CDLL(ctypes.util.find_library("c") or "libc.so.6", use_errno=True)
    libc.mmap.restype = C.c_void_p
    libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
    libc.syscall.restype = C.c_long
    size = (nbytes + (1 << 21) - 1) & ~((1 << 21) - 1)
    addr = libc.mmap(None, size, _PROT_RW, _MAP_PRIVATE_ANON, -1, 0)
    if addr in (None, C.c_void_p(-1).value):
        raise OSError("mmap failed")
    libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    libc.madvise(addr, size, _MADV_HUGEPAGE)  # best effort
    mask = (C.c_ulong * 16)()
    mask[node // 64] = 1 << (node % 64)
    rc = libc.syscall(C.c_long(_SYS_MBIND), C.c_void_p(addr), C.c_ulong(size), C.c_int(_MPOL_BIND), mask, C.c_ulong(16 * 64), C.c_uint(0))
    if rc != 0:
        libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
        libc.munmap(addr, size)
        raise OSError(C.get_errno(), "mbind failed")
    C.memset(addr, 0, size)  # first touch under the policy
    return addr, size


def pinned_empty(shape, dtype, device_index: int = 0):
    """A page-locked CPU tensor of `shape`/`dtype` for DMA with GPU `device_index` (contents: zeros or garbage)."""
    import torch

    numel = 1
    for d in shape:
        numel *= int(d)
    nbytes = max(numel * torch.empty((), dtype=dtype).element_size(), 1)
    node = gpu_numa_node(device_index) if os.environ.get("EVG_HOST_NUMA", "1") != "0" else None
    if node is not None and _numa_nodes_online() > 1 and nbytes >= (1 << 20):
        try:
            addr, size = _bound_mapping(nbytes, node)
            rc = torch.cuda.cudart().cudaHostRegister(addr, size, 0)
            if int(rc) != 0:
                raise OSError("cudaHostRegister returned %s" % rc)
            _keep.append((addr, size))
            buf = (C.c_char * nbytes).from_address(addr)
            t = torch.frombuffer(buf, dtype=dtype, count=numel).reshape(shape)
            _last.update(method="mbind+register", numa_node=node)
            return t
        except Exception:
            pass
    _last.update(method="torch", numa_node=node)
    return torch.empty(shape, dtype=dtype).pin_memory()
