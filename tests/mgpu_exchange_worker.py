"""Worker of tests/test_gpu_exchange.py::test_exchange_over_nccl (launched by torch.distributed.run, one rank per GPU): the
real transport -- torch.distributed NCCL all-to-all over NVLink -- under the same stream / oracle comparison as the
single-GPU LocalTransport test."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from denormalized_b200 import GpuStreamingWindow, canonical_schema
    from denormalized_b200.exchange import TorchTransport, exchange_step
    from tests.helpers import DEFAULT_AGGS, assert_rows_equal, record_batch_rows, rows_to_batch, run_oracle_batches, to_record_batch
    from tests.test_gpu_exchange import T0, _stream
    L, S, filt = 4000, 1000, ("max", ">", 100)
    rng = np.random.default_rng(5)
    batches, t_end = _stream(rng, 40, 2000, 900, 200, long_keys=True)          # identical on every rank (same seed)
    close = (t_end // 1000 + 1) * 1000 + 2 * L
    w = GpuStreamingWindow(canonical_schema(), "sensor_name", DEFAULT_AGGS, L, S, filt, device=local, expected_groups=128)
    w.set_exchange(rank, world)
    tr = TorchTransport(device=f"cuda:{local}")
    got = []
    for i, b in enumerate(batches):
        if i % world == rank:
            w.push(to_record_batch(b))
        if i % 8 == 7:
            got += record_batch_rows(exchange_step(w, tr))
    w.push(to_record_batch(rows_to_batch([(close, 1.0, b"sentinel")])))
    got += record_batch_rows(exchange_step(w, tr))
    st = w.stats()
    w.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (got, st["exchanged_out"], st["exchanged_in"]))
    want = None
    if rank == 0:
        want = run_oracle_batches(batches + [rows_to_batch([(close, 1.0, b"sentinel")])], L, S, filt)
        rows = [r for g in gathered for r in g[0]]
        assert len(want) > 500 and all(len(g[0]) > 0 for g in gathered)
        assert sum(g[1] for g in gathered) == sum(g[2] for g in gathered) > 1000
        assert_rows_equal(rows, want)
        print(f"NCCL exchange ok: {len(rows)} rows from {world} ranks match the oracle; packets {sum(g[1] for g in gathered)}")

    # ---- the fused exchange (dnz_group): CUDA IPC mappings of the peers' receive rings, P2P stores over NVLink, interprocess
    # events; torch.distributed is only the rendezvous helper at creation
    from denormalized_b200 import ExchangeGroup

    def rendezvous(blob):
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(out, t)
        raw = out.cpu().numpy().tobytes()
        return [raw[i * len(blob):(i + 1) * len(blob)] for i in range(world)]
    grp = ExchangeGroup.create(rank, world, local, rendezvous, ring_entries=1 << 16, ring_key_bytes=4 << 20)
    w = GpuStreamingWindow(canonical_schema(), "sensor_name", DEFAULT_AGGS, L, S, filt, device=local, expected_groups=2048)
    grp.attach(w)
    got = []
    for i, b in enumerate(batches):
        if i % world == rank:
            w.push(to_record_batch(b))
        if i % 8 == 7:
            grp.step(w)
            got += record_batch_rows(w.poll())
    w.push(to_record_batch(rows_to_batch([(close, 1.0, b"sentinel")])))
    grp.flush(w)                                  # end of stream: process + three steps (publish | pack | merge + emit)
    got += record_batch_rows(w.poll())
    grp.step(w)                                   # one more (empty) step makes the packet counters visible
    st = w.stats()
    w.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (got, st["exchanged_out"], st["exchanged_in"]))
    if rank == 0:
        rows = [r for g in gathered for r in g[0]]
        assert all(len(g[0]) > 0 for g in gathered)
        assert sum(g[1] for g in gathered) == sum(g[2] for g in gathered) > 1000
        assert_rows_equal(rows, want)
        print(f"fused P2P exchange ok: {len(rows)} rows from {world} ranks match the oracle; packets {sum(g[1] for g in gathered)}")
    dist.barrier()
    grp.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
