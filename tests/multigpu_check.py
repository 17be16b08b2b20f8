"""Run under torchrun (one rank per GPU): the in-library sharded search (csrc/shard.cu: NCCL all-gathers, CUDA-graph
replay) against the unsharded engine on the same device -- bit-identical ids / distances / counts -- and against
the CPU oracle.  `python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py`"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ivfadc_jl_b200 as iv
from ivfadc_jl_b200 import sharded, synth
from oracle import oracle as orc

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
ok = True
for (D, m, kc, n, nq, k, w) in ((128, 16, 64, 60000, 1003, 10, 16), (96, 12, 40, 30000, 257, 5, 8), (64, 8, 16, 5000, 31, 10, 4)):
    X = synth.blobs(n, D, kc, seed=11); Q = synth.blobs(nq, D, kc, seed=12)
    cent, cb, codes = synth.random_quantizers(kc, D, m, 256, seed=13, data=X)
    qz = orc.Quantizers(cent, cb, codes)
    full = iv.IVFADCIndex.from_quantizers(cent, cb, codes, device=lr)            # unsharded twin on this device
    cells = full._add(X, iv._capi.LAST, want_cells=True)
    owners = sharded.balanced_owners(np.bincount(cells, minlength=kc), world)
    e = iv.IVFADCIndex.from_quantizers(cent, cb, codes, device=lr, shard=(rank, world))
    e.set_cell_owners(owners)
    e._add(X, iv._capi.LAST, assign=cells.astype(np.int64), assign_base=0)
    assert int(e.list_sizes().sum()) == int(np.bincount(cells, minlength=kc)[owners == rank].sum())
    sharded.init_comm(e)
    ui, ud, uc = full.search_packed(Q, k, w)
    dQ = torch.from_numpy(Q).to(dev)
    for it in range(5):   # eager, eager, capture, replay, replay
        gi, gd, gc = sharded.search_sharded_device(e, dQ, k, w)
        e.check_async(torch.cuda.current_stream().cuda_stream)
        gi, gd, gc = gi.cpu().numpy().view(np.uint64), gd.cpu().numpy(), gc.cpu().numpy()
        same = np.array_equal(gc, uc) and np.array_equal(gi, ui) and np.array_equal(gd.view(np.uint8), ud.view(np.uint8))
        if not same:
            ok = False
            print(f"rank {rank}: device variant differs from the unsharded engine (shape {D},{m},{kc}; call {it})", flush=True)
    for it in range(4):
        hi, hd, hc = sharded.search_sharded_host(e, Q, k, w)
        same = np.array_equal(hc, uc) and np.array_equal(hi, ui) and np.array_equal(hd.view(np.uint8), ud.view(np.uint8))
        if not same:
            ok = False
            print(f"rank {rank}: host variant differs from the unsharded engine (shape {D},{m},{kc}; call {it})", flush=True)
    if rank == 0:
        order = np.argsort(cells, kind="stable")
        off = np.zeros(kc + 1, dtype=np.int64); np.cumsum(np.bincount(cells, minlength=kc), out=off[1:])
        _, ocodes = orc.encode(qz, X, nthreads=8)
        oi, od, oc, _ = orc.search_csr(qz, off, ocodes[order], order.astype(np.uint64), Q, k, w, nthreads=8)
        rep = orc.compare_search(hi, hd, hc, oi, od, oc, rtol=1e-5)
        print(f"shape D={D} m={m} kc={kc}: sharded == unsharded bits: {ok}; vs oracle max rel err {rep['max_rel_err']:.2e}, "
              f"{rep['near_tie_id_mismatches']} near-tie swaps; stats {e.stats()}", flush=True)
    torch.cuda.synchronize(); dist.barrier()
    e.close(); full.close()
t = torch.tensor([0 if ok else 1], device=dev); dist.all_reduce(t)
if rank == 0:
    print("MULTIGPU_CHECK", "OK" if int(t.item()) == 0 else "FAILED", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 0 else 1)
