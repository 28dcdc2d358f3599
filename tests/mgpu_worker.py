"""Worker of tests/test_gpu_multi.py (launched with torchrun, one process per GPU): source-sharded align of the
golden pair and of a synthetic pair, checked against the oracle on rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle
    import toyslam_b200 as nb
    from toyslam_b200.sharding import ShardedNdt
    from util import load_pair, rel_err, transform_delta

    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    out = {"rank": rank, "world": world, "cases": []}
    for name, method, kw in (("pair_ds0p1.npz", oracle.DIRECT7, {}), ("pair_ds0p3.npz", oracle.DIRECT7, {"eps": 0.01, "max_iter": 64}),
                             ("pair_ds0p1.npz", oracle.DIRECT26, {})):
        tgt, src = load_pair(name)
        ndt = nb.NormalDistributionsTransform(device=local)
        ndt.setNeighborhoodSearchMethod(method)
        if "eps" in kw:
            ndt.setTransformationEpsilon(kw["eps"]); ndt.setMaximumIterations(kw["max_iter"])
        sh = ShardedNdt(ndt, dist)
        sh.setInputTarget(tgt)
        sh.setInputSource(src)
        e = sh.eval_derivatives(np.array([0.3, 0.1, -0.02, 0.004, -0.002, -0.01]))
        sh.align()
        r = sh.result()
        fit = sh.getFitnessScore()
        # every rank must hold identical bits (rank-ordered sum, redundant identical Newton step)
        blob = [None] * world
        dist.all_gather_object(blob, (r["final"].tobytes(), r["iterations"], r["n_evaluations"], float(e["score"])))
        same = all(b == blob[0] for b in blob)
        case = {"name": name, "method": method, "identical_across_ranks": same}
        if rank == 0:
            ref = oracle.NormalDistributionsTransform()
            ref.setNeighborhoodSearchMethod(method)
            if "eps" in kw:
                ref.setTransformationEpsilon(kw["eps"]); ref.setMaximumIterations(kw["max_iter"])
            ref.setInputTarget(tgt); ref.setInputSource(src)
            eo = ref.eval_derivatives(np.array([0.3, 0.1, -0.02, 0.004, -0.002, -0.01]))
            ref.align()
            rr = ref.result()
            dt, dr = transform_delta(r["final"], rr["final"])
            case.update({"fitness_rel": abs(fit - ref.getFitnessScore()) / ref.getFitnessScore()})
            case.update({"hits_equal": int(e["hits"]) == int(eo["hits"]), "score_rel": abs(e["score"] - eo["score"]) / abs(eo["score"]),
                         "grad_rel": rel_err(e["gradient"], eo["gradient"]), "hess_rel": rel_err(e["hessian"], eo["hessian"]),
                         "iterations": [r["iterations"], rr["iterations"]], "evaluations": [r["n_evaluations"], rr["n_evaluations"]],
                         "hessian_passes": [r["n_hessian_passes"], rr["n_hessian_passes"]], "dt": dt, "dr": dr,
                         "tp_rel": abs(r["trans_probability"] - rr["trans_probability"]) / abs(rr["trans_probability"])})
        out["cases"].append(case)
        ndt.comm_detach()
        dist.barrier()
    # ---- sharded target-map build: every rank reduces its point range, partials exchanged, merged in rank order ----
    from util import synthetic_scene
    out["build"] = []
    for name, res in (("pair_ds0p1.npz", 1.0), ("pair_ds0p3.npz", 2.0), ("synthetic_offset", 0.5)):
        if name == "synthetic_offset":
            tgt, src = synthetic_scene(offset=(2000.0, -1500.0, 50.0), seed=4)
        else:
            tgt, src = load_pair(name)
        ndt = nb.NormalDistributionsTransform(device=local)
        ndt.setResolution(res)
        sh = ShardedNdt(ndt, dist)
        st = sh.setInputTargetSharded(tgt)
        info = ndt.map_info()
        g = ndt.dump_voxels()
        blob = [None] * world
        dist.all_gather_object(blob, (g["keys"].tobytes(), g["counts"].tobytes(), g["mean"].tobytes(), g["icov"].tobytes()))
        case = {"name": name, "status": int(st), "identical_across_ranks": all(b == blob[0] for b in blob)}
        if rank == 0:
            ref = oracle.NormalDistributionsTransform()
            ref.setResolution(res)
            ref.setInputTarget(tgt)
            rl = ref.dump_leaves()
            single = nb.NormalDistributionsTransform(device=local)
            single.setResolution(res)
            single.setInputTarget(tgt)
            s1 = single.dump_voxels()
            case.update({"keys_equal": bool(np.array_equal(g["keys"], rl["keys"])), "counts_equal": bool(np.array_equal(g["counts"], rl["counts"])),
                         "mean_rel": rel_err(g["mean"], rl["mean"]),
                         "icov_rel_vs_single_gpu": float(np.max(np.abs(g["icov"] - s1["icov"]) / (np.abs(s1["icov"]).max(axis=(1, 2), keepdims=True) + 1e-300))),
                         "n_voxels": [int(info["n_voxels"]), int(len(rl["keys"]))], "n_valid": int(info["n_valid"])})
            # the merged map must drive the solver like the single-GPU map
            ndt.setInputSource(src); single.setInputSource(src)
            p = np.array([0.2, -0.1, 0.02, 0.003, -0.002, 0.01]) + (np.array([2000.0, -1500.0, 50.0, 0, 0, 0]) if name == "synthetic_offset" else 0)
            a, b = ndt.eval_derivatives(p), single.eval_derivatives(p)
            case.update({"hits_equal": int(a["hits"]) == int(b["hits"]), "grad_rel": rel_err(a["gradient"], b["gradient"]),
                         "hess_rel": rel_err(a["hessian"], b["hessian"])})
        out["build"].append(case)
        dist.barrier()
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
