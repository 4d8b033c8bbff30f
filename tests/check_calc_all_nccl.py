"""Multi-GPU check of the public sharded call, run under torch.distributed.run on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tests/check_calc_all_nccl.py

Every rank builds the same generator (seeded synthetic inputs, host handles and device-resident handles), calls
gen.calc_all(dst=0) / calc_all(dst=None) over NCCL and calc_to_file(shard=True) into one shared .npy; rank 0 compares
everything with its own serial per-timeslice results and with the oracle.  Prints one JSON line on rank 0."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import easydistillation_b200 as edb  # noqa: E402
from oracle import elemental_oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {"world": world}
    # Lt not a multiple of the world size or of the chunk: uneven ranges, partial last chunks
    for latt, Ne, nabla, nmom, chunk in [([8, 4, 4, 11], 12, 1, 7, 2), ([16, 8, 8, 2 * world + 1], 40, 2, 33, 4)]:
        Lt = latt[3]
        moms = orc.momentum_set(nmom)
        U = np.stack([orc.synthetic_links(latt, t) for t in range(Lt)])
        V = np.stack([orc.synthetic_eigvecs(latt, Ne, t) for t in range(Lt)])
        gen = edb.ElementalGenerator(latt, edb.GaugeFieldHostmem(U), edb.EigenvectorHostmem(V), nabla, moms, device=local)
        gen.load("x")
        got = gen.calc_all(dst=0, chunk=chunk)
        everywhere = gen.calc_all(dst=None, chunk=chunk)
        assert (got is not None) == (rank == 0)
        gdev = edb.ElementalGenerator(latt, edb.GaugeFieldDevice([torch.from_numpy(U[t]).to(dev) for t in range(Lt)]),
                                      edb.EigenvectorDevice([torch.from_numpy(V[t]).to(dev) for t in range(Lt)]), nabla, moms, device=local)
        gdev.load("x")
        got_dev = gdev.calc_all(dst=0, chunk=chunk)
        tmp = [tempfile.mkdtemp(prefix="edk_nccl_") if rank == 0 else None]
        dist.broadcast_object_list(tmp, src=0)
        handle = edb.ElementalNpy(os.path.join(tmp[0], "cfg_"), ".elemental.npy", [4 if nabla == 1 else 13, len(moms), Lt, Ne, Ne], Ne)
        gen.calc_to_file(handle, "a", shard=True)
        dist.barrier()
        serial = torch.stack([gen.calc_device(t).clone() for t in range(Lt)])
        err_all = float((everywhere - serial).abs().max())  # every rank holds the full result
        t = torch.tensor([err_all], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            ref = np.stack([orc.elemental_timeslice_closed_form(V[t_], orc.links_file_to_spatial(U[t_]), latt, nabla, moms) for t_ in (0, Lt - 1)])
            g = got.cpu().numpy()
            worst = max(float(np.linalg.norm(g[t_] - r) / np.linalg.norm(r)) for t_, r in zip((0, Lt - 1), ref))
            filed = np.load(os.path.join(tmp[0], "cfg_a.elemental.npy"), mmap_mode="r")
            report[f"{latt}"] = {
                "bit_identical_to_serial": bool(torch.equal(got, serial)),
                "device_handles_bit_identical": bool(torch.equal(got_dev, serial)),
                "dst_none_max_abs_diff_over_ranks": float(t.item()),
                "file_equals_result": bool(np.array_equal(np.asarray(filed), np.transpose(g, (1, 2, 0, 3, 4)))),
                "rel_err_vs_oracle": worst,
            }
            assert report[f"{latt}"]["bit_identical_to_serial"] and report[f"{latt}"]["device_handles_bit_identical"]
            assert report[f"{latt}"]["file_equals_result"] and t.item() == 0.0 and worst < 1e-10
        dist.barrier()
    if rank == 0:
        report["ok"] = True
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
