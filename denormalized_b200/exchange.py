"""Multi-GPU pane exchange driver: the one collective of the path (include/dnz_gpu.h, "multi-GPU pane exchange").

Replaces RepartitionExec(Hash(group keys)) (physical_optimizer/coalesce_before_streaming_window_aggregate.rs:63-73) for input
that is NOT key-partitioned: every rank aggregates the batches it was dealt; `exchange_step` then runs, collectively,

    local watermark -> all-reduce(min) -> export partial pane states by owner -> ONE all-to-all of packets (+ one of key
    bytes; their split sizes travel in a small all-to-all before) -> owner merge -> every rank emits its own keys.

`Transport` hides where the bytes move: `TorchTransport` uses torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU
tests, where the "device" buffers are host arrays), `LocalTransport` moves packets between several operators of ONE process
(single-GPU tests).  The packet format is PartialEntry (64 B, denormalized_b200/csrc/dnz_kernels.h)."""
from __future__ import annotations

import numpy as np

PARTIAL_BYTES = 64
PARTIAL_DTYPE = np.dtype([("pane", "<i8"), ("cnt", "<u8"), ("sum", "<f8"), ("minkey", "<u8"), ("maxkey", "<u8"),
                          ("nullrows", "<u8"), ("fz", "<u8"), ("key_off", "<u4"), ("key_len", "<u4")])
assert PARTIAL_DTYPE.itemsize == PARTIAL_BYTES
INT64_MIN = -(2 ** 63)


def plan_splits(owner_counts, owner_key_bytes):
    """Send-side split sizes in BYTES for the two payload all-to-alls, from the per-owner packet / key-byte counts."""
    return [int(c) * PARTIAL_BYTES for c in owner_counts], [int(b) for b in owner_key_bytes]


class TorchTransport:
    """torch.distributed collectives on uint8 tensors.  `device` is where the packet buffers live ("cuda:N" for the real
    operator; "cpu" for the gloo tests)."""

    def __init__(self, group=None, device="cuda"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def global_watermark(self, local_wm):
        t = self.torch.tensor([INT64_MIN if local_wm is None else int(local_wm)], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        v = int(t.item())
        return None if v == INT64_MIN else v

    def exchange_sizes(self, rows):
        """rows: world x k int64 (what this rank sends to each peer) -> world x k (what each peer sends to this rank)."""
        t = self.torch.tensor(rows, dtype=self.torch.int64, device=self.device).reshape(self.world, -1).contiguous()
        out = self.torch.empty_like(t)
        self.dist.all_to_all_single(out, t, group=self.group)
        return out.cpu().numpy()

    def all_to_all_bytes(self, send, send_splits, recv_splits):
        """send: uint8 tensor on self.device; returns the received uint8 tensor (concatenated in rank order)."""
        recv = self.torch.empty(int(sum(recv_splits)), dtype=self.torch.uint8, device=self.device)
        self.dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in recv_splits],
                                    input_split_sizes=[int(x) for x in send_splits], group=self.group)
        return recv

    def pane_range(self, lo, hi):
        t = self.torch.tensor([lo, -hi], dtype=self.torch.int64, device=self.device)       # min(lo), max(hi)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return int(t[0].item()), -int(t[1].item())


class _DevView:
    """Zero-copy torch view of `nbytes` of device memory owned by the operator (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def device_bytes(torch, ptr, nbytes, device):
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    if torch.device(device).type == "cpu":          # gloo tests: the "device" buffers are host arrays
        import ctypes
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * int(nbytes)).from_address(int(ptr))))
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


def exchange_step(window, transport, emit=True):
    """One collective exchange step of a GpuStreamingWindow in exchange mode.  Returns the emitted RecordBatch (this rank's
    keys, every window that closed under the GLOBAL watermark); emit="device" returns the device-resident result struct of
    dnz_window_poll_device instead; emit=False only moves and merges the partials."""
    torch = transport.torch
    gw = transport.global_watermark(window.process())
    parts = window.export_partials(gw)
    send_e, send_k = plan_splits(parts["owner_counts"], parts["owner_key_bytes"])
    sizes = transport.exchange_sizes([[parts["owner_counts"][o], parts["owner_key_bytes"][o]] for o in range(transport.world)])
    src_counts, src_kbytes = [int(x) for x in sizes[:, 0]], [int(x) for x in sizes[:, 1]]
    lo, hi = parts["pane_lo"], parts["pane_hi"]
    glo, ghi = transport.pane_range(lo if hi >= lo else 2 ** 62, hi if hi >= lo else -(2 ** 62))
    ent = device_bytes(torch, parts["entries"][0], parts["entries"][1] * PARTIAL_BYTES, transport.device)
    keys = device_bytes(torch, parts["keys"][0], parts["keys"][1], transport.device)
    recv_e = transport.all_to_all_bytes(ent, send_e, [c * PARTIAL_BYTES for c in src_counts])
    recv_k = transport.all_to_all_bytes(keys, send_k, src_kbytes)
    if sum(src_counts):
        if transport.device.type == "cuda":
            torch.cuda.synchronize(transport.device)    # the packets are consumed on the operator's own stream
        window.import_partials(recv_e.data_ptr(), src_counts, recv_k.data_ptr() if recv_k.numel() else 0, src_kbytes, glo, ghi)
    if not emit:
        return None
    if gw is not None:
        window.flush(gw)
    return window.poll_device() if emit == "device" else window.poll()


class LocalTransport:
    """All ranks live in this process (one operator each, same GPU): `exchange_all` performs what `exchange_step` does on
    every rank, moving the packets with device-to-device tensor copies.  Used by the single-GPU tests."""

    def __init__(self, windows, device="cuda:0"):
        import torch
        self.torch, self.windows, self.world, self.device = torch, windows, len(windows), torch.device(device)

    def exchange_all(self):
        torch = self.torch
        lws = [w.process() for w in self.windows]
        gw = None if any(x is None for x in lws) else min(lws)
        parts = [w.export_partials(gw) for w in self.windows]
        ranges = [(p["pane_lo"], p["pane_hi"]) for p in parts if p["pane_hi"] >= p["pane_lo"]]
        glo, ghi = (min(r[0] for r in ranges), max(r[1] for r in ranges)) if ranges else (0, -1)
        out = []
        for dst, w in enumerate(self.windows):
            e_chunks, k_chunks, src_counts, src_kbytes = [], [], [], []
            for src, p in enumerate(parts):
                e0 = sum(p["owner_counts"][:dst]) * PARTIAL_BYTES; k0 = sum(p["owner_key_bytes"][:dst])
                ne, nk = p["owner_counts"][dst] * PARTIAL_BYTES, p["owner_key_bytes"][dst]
                ent = device_bytes(torch, p["entries"][0], p["entries"][1] * PARTIAL_BYTES, self.device)
                keys = device_bytes(torch, p["keys"][0], p["keys"][1], self.device)
                e_chunks.append(ent[e0:e0 + ne]); k_chunks.append(keys[k0:k0 + nk])
                src_counts.append(p["owner_counts"][dst]); src_kbytes.append(nk)
            recv_e, recv_k = torch.cat(e_chunks).contiguous(), torch.cat(k_chunks).contiguous()
            out.append((recv_e, recv_k, src_counts, src_kbytes))
        torch.cuda.synchronize(self.device)
        results = []
        for w, (recv_e, recv_k, sc, sk) in zip(self.windows, out):
            if sum(sc):
                w.import_partials(recv_e.data_ptr(), sc, recv_k.data_ptr() if recv_k.numel() else 0, sk, glo, ghi)
            if gw is not None:
                w.flush(gw)
            results.append(w.poll())
        return results
