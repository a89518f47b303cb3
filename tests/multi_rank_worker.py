"""Worker of tests/test_multi_gpu.py::test_one_process_per_gpu (launched by torch.distributed.run, one rank per GPU).
Every rank joins the device group with rtb200_create_rank, runs the host-buffer and the resident calls collectively and
compares ITS slab of the results with the single-device result it computes on its own GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import radiativetransfer_b200 as rt
    from conftest import rel_err
    from radiativetransfer_b200 import workloads as W

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")          # plumbing only: broadcasts the NCCL unique ids of the library's own groups

    def new_uid():                            # every communicator needs a fresh id
        uid = [rt.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        return uid[0]

    bg = W.uvb_background(3.0)
    ksi = np.concatenate([bg["ksi24"], bg["ksi25"], bg["ksi26"]])
    spectra = W.synthetic_spectra()
    for g in (W.uniform_grid(24, seed=3), W.nested_grid(6, 2, W.central_box_refine(0.2, 0.7, levels=2), seed=5)):
        N = g["level"].size
        one = rt.Transport(device=local)
        one.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
        J1, nseg1 = one.diffuse(bg["uvb"], bg["beta"])
        src = np.array([N // 2, N // 3, 5, N - 7], dtype=np.int32); wt = np.array([1, 2, 1, 3], dtype=np.int32)
        p1 = one.point(spectra, src, wt)
        one.close()
        for mode in (1, 0):
            grp = rt.Transport(device=local, comm=(world, rank, new_uid()))
            grp.set_tuning(multi_reduce=mode)
            grp.set_grid(g["nx"], g["level"], g["HI"], g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
            info = grp.info()
            assert info["nranks"] == world and info["nlocal"] == 1 and info["first_rank"] == rank
            if mode == 1 and info["reduce_mode"] != 1:
                print(f"rank {rank}: cudaIpc mapping unavailable here, NCCL reduce-scatter used instead")
            off, cnt, _, _, _ = grp.slab(0)
            # host-buffer call: every rank reads / writes only its slab of the caller's arrays
            J = np.full((3, N), -1.0)
            _, nseg = grp.diffuse(bg["uvb"], bg["beta"], out=J)
            t = torch.tensor([nseg], dtype=torch.int64); dist.all_reduce(t)
            assert int(t[0]) == nseg1
            assert rel_err(J[:, off:off + cnt], J1[:, off:off + cnt]) < 1e-13
            outside = np.ones(N, dtype=bool); outside[off:off + cnt] = False
            assert np.all(J[:, outside] == -1.0)
            # species: slab upload + all-gather gives every rank the whole array
            hi = g["HI"] * 0.5
            mine = np.zeros(N); mine[off:off + cnt] = hi[off:off + cnt]       # a rank only owns its slab
            grp.update_species(HI=mine)
            grp.diffuse_resident(bg["uvb"], bg["beta"], ksi=ksi)
            grp.sync()
            _, _, Js, Ks, _ = grp.slab_get(0)
            ref = rt.Transport(device=local)
            ref.set_grid(g["nx"], g["level"], hi, g["HeI"], g["HeII"], g["rho"], g["abun2"], g["box_size"])
            Jr, _ = ref.diffuse(bg["uvb"], bg["beta"])
            ref.close()
            assert rel_err(Js, Jr[:, off:off + cnt]) < 1e-13
            grp.update_species(HI=g["HI"])
            # point sources
            p = grp.point(spectra, src, wt)
            assert rel_err(p["rates"][:, off:off + cnt], p1["rates"][:, off:off + cnt], floor=1e-300) < 1e-12
            mine_src = np.arange(rank, src.size, world)
            if mine_src.size:       # fewer sources than ranks: this rank casts none
                assert rel_err(p["ndot_remaining"][mine_src], p1["ndot_remaining"][mine_src], floor=1e-300) < 1e-12
            grp.close()
    dist.barrier()
    print("RANK-OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
