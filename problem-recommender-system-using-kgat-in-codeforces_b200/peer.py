"""NVLink peer memory for the row-sharded propagation (csrc/peer.cu): one IPC-exportable allocation per rank
("arena") carved into named fp32 tables plus a flag pad, with every other rank's arena mapped into this process.

``torch.distributed`` is used once, to exchange the 64-byte IPC handles; after that every exchange is stores into
peer memory + a flag handshake, all stream-ordered device work that CUDA graphs capture.
"""

from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from ._lib import KgatLibraryError, check

_ALIGN = 64  # floats (256 B)


class _RawCuda:
    """Minimal ``__cuda_array_interface__`` carrier so torch can view memory it did not allocate."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerArena:
    def __init__(self, rank: int, world: int, device, tables: dict[str, tuple[int, int]], n_channels: int, timeout_s: float = 10.0):
        """``tables``: name -> (rows, d).  Every rank must pass the same specification."""
        if world - 1 > 31:
            raise KgatLibraryError("PeerArena supports at most 32 ranks (one node)")
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.peers = [q for q in range(world) if q != rank]
        self.n_channels = n_channels
        self.lib = _lib.load()
        off = 0
        self._off: dict[str, int] = {}
        self._shape = dict(tables)
        for name, (rows, d) in tables.items():
            self._off[name] = off
            off += (rows * d + _ALIGN - 1) // _ALIGN * _ALIGN
        self._flag_off = off  # in floats == int32 slots
        n_flag = (n_channels * max(world - 1, 1) + _ALIGN - 1) // _ALIGN * _ALIGN
        self.n_bytes = 4 * (off + n_flag)
        base = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.kgat_peer_alloc(self.n_bytes, C.byref(base)), "peer_alloc")
            self.base = int(base.value)
            handle = C.create_string_buffer(64)
            check(self.lib.kgat_peer_export(self.base, handle), "peer_export")
            handles: list = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw))
            self.peer_base: dict[int, int] = {}
            for q in self.peers:
                p = C.c_void_p()
                check(self.lib.kgat_peer_import(C.create_string_buffer(handles[q], 64), C.byref(p)), f"peer_import(rank {q})")
                self.peer_base[q] = int(p.value)
        self._views = {name: torch.as_tensor(_RawCuda(self.base + 4 * self._off[name], shape, "<f4"), device=self.device)
                       for name, shape in self._shape.items()}
        self.flags = torch.as_tensor(_RawCuda(self.base + 4 * self._flag_off, (n_channels, max(world - 1, 1)), "<i4"), device=self.device)
        self.seq = torch.zeros(n_channels, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        clock_khz = getattr(torch.cuda.get_device_properties(self.device), "clock_rate", 1_900_000)
        self.timeout_cycles = int(timeout_s * clock_khz * 1e3)
        # my slot in peer q's flag pad: peers are numbered by rank with the owner left out
        self._flag_ptrs = []
        for c in range(n_channels):
            ptrs = [self.peer_base[q] + 4 * (self._flag_off + c * (world - 1) + (rank if rank < q else rank - 1)) for q in self.peers]
            self._flag_ptrs.append(torch.tensor(ptrs or [0], dtype=torch.int64, device=self.device))
        self._ptr_cache: dict = {}
        dist.barrier()  # nobody stores into a peer before every mapping exists

    # ------------------------------------------------------------------------------------------
    def table(self, name: str) -> torch.Tensor:
        return self._views[name]

    def peer_ptrs(self, name: str, row_offset: int = 0) -> torch.Tensor:
        """Device array (int64) of pointers to row ``row_offset`` of table ``name`` in every peer's arena."""
        key = (name, row_offset)
        if key not in self._ptr_cache:
            d = self._shape[name][1]
            ptrs = [self.peer_base[q] + 4 * (self._off[name] + row_offset * d) for q in self.peers]
            self._ptr_cache[key] = torch.tensor(ptrs or [0], dtype=torch.int64, device=self.device)
        return self._ptr_cache[key]

    def push(self, name: str, row_offset: int, n_rows: int, max_ctas: int = 0) -> None:
        """Copy rows [row_offset, row_offset + n_rows) of my table to the same rows of every peer's table."""
        if not self.peers:
            return
        t = self._views[name]
        src = t[row_offset : row_offset + n_rows]
        check(self.lib.kgat_peer_push(src.data_ptr(), self.peer_ptrs(name, row_offset).data_ptr(), len(self.peers), src.numel(),
                                      int(max_ctas), torch.cuda.current_stream().cuda_stream), "peer_push")

    def push_rows(self, name: str, rows: torch.Tensor, count_dev: torch.Tensor, max_rows: int) -> None:
        """Copy the listed rows (int32 node ids, device-side count) of my table to the same rows of every peer's table."""
        if not self.peers:
            return
        t = self._views[name]
        check(self.lib.kgat_peer_push_rows(t.data_ptr(), self.peer_ptrs(name, 0).data_ptr(), len(self.peers), rows.data_ptr(), count_dev.data_ptr(),
                                           int(max_rows), t.shape[1], t.stride(0), torch.cuda.current_stream().cuda_stream), "peer_push_rows")

    def copy(self, name: str, row_offset: int, n_rows: int) -> None:
        """Same transfer as ``push`` on the copy engines (one cudaMemcpyAsync per peer): no SM involved."""
        d = self._shape[name][1]
        off = 4 * (self._off[name] + row_offset * d)
        stream = torch.cuda.current_stream().cuda_stream
        for q in self.peers:
            check(self.lib.kgat_peer_copy(self.peer_base[q] + off, self.base + off, 4 * n_rows * d, stream), "peer_copy")

    def signal_wait(self, channel: int) -> None:
        """All peers' stores of this channel have landed here once this (stream-ordered) call has run."""
        if not self.peers:
            return
        check(self.lib.kgat_peer_signal_wait(self._flag_ptrs[channel].data_ptr(), self.flags[channel].data_ptr(), len(self.peers),
                                             self.seq[channel :].data_ptr(), self.status.data_ptr(), self.timeout_cycles,
                                             torch.cuda.current_stream().cuda_stream), "peer_signal_wait")

    def check(self) -> None:
        if int(self.status.item()) != 0:
            raise KgatLibraryError(f"rank {self.rank}: a peer did not arrive at a row exchange within the time-out")

    def close(self) -> None:
        if self.base is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier()
        self._views.clear()
        with torch.cuda.device(self.device):
            for p in self.peer_base.values():
                self.lib.kgat_peer_close(p)
            self.lib.kgat_peer_free(self.base)
        self.base = None
