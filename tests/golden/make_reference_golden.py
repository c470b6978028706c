"""Generates tests/golden/reference_outputs.npz with THE REFERENCE'S OWN CODE: oracle/_ref/libsrl_reference.so, i.e.
/root/reference/src/{optimize,utility,eskfEstimator,cloudMap,state,lioOptimization,rgbMapTracker}.cpp compiled unmodified where
they lie over the stand-in headers of oracle/shim/ (DESIGN.md section 2).  Inputs are the ones of scan_matching.npz / sweep_prep.npz /
map_insert.npz, so the fixture holds only outputs.  These vectors travel to machines without /root/reference (the GPU box) and pin
both the oracle (tests/test_golden.py, CPU) and the CUDA path (GPU) to what the reference's functions returned here.
The arithmetic inside Eigen calls is the shim's restatement (unpinned).  Run from the repo root, where /root/reference
exists:  python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O          # noqa: E402  (parameter structs only)
from oracle import reference_py as Rf      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BIG = 2 ** 31 - 1
PASS_CASES = {"nb1": dict(max_num_residuals=BIG, frame_id=100), "nb2": dict(max_num_residuals=BIG, frame_id=5),
              "cap": dict(max_num_residuals=100, frame_id=100), "cap600": dict(max_num_residuals=600, frame_id=100),
              "dflt": dict(max_num_residuals=-1, frame_id=100)}
IEKF_CASES = {"steady": dict(max_num_residuals=BIG), "cap600": dict(max_num_residuals=600), "init": dict(max_num_residuals=BIG, frame_id=5)}


def main():
    sm = np.load(os.path.join(HERE, "scan_matching.npz"))
    sp = np.load(os.path.join(HERE, "sweep_prep.npz"))
    mi = np.load(os.path.join(HERE, "map_insert.npz"))
    out = {}
    ref = Rf.Reference()
    ref.load(sm["map_keys"], sm["map_counts"], sm["map_xyz"])
    for tag, kw in PASS_CASES.items():                                  # lioOptimization::buildPlaneResiduals
        r = ref.build_plane_residuals(sm["raw_xyz"], sm["q_init"], sm["t_init"], sm["t_last"], O.r3live_params(**kw))
        assert not r["threw"]
        out.update({f"pass_{tag}_rows": r["rows"], f"pass_{tag}_world": r["world_xyz"], f"pass_{tag}_loss": np.array(r["loss_sum"]),
                    f"pass_{tag}_success": np.array(r["success"]), f"pass_{tag}_num_residuals": np.array(r["num_residuals_used"])})
    for tag, kw in IEKF_CASES.items():                                  # lioOptimization::updateIEKF
        e0 = O.Eskf(p=sm["t_init"].copy(), q=sm["q_init"].copy(), cov=sm["prior_cov"].copy())
        r = ref.update_iekf(sm["raw_xyz"], e0, sm["t_last"], O.r3live_params(**kw))
        assert not r["threw"] and r["success"]
        e = r["eskf"]
        out.update({f"iekf_{tag}_p": e.p, f"iekf_{tag}_q": e.q, f"iekf_{tag}_v": e.v, f"iekf_{tag}_ba": e.ba, f"iekf_{tag}_bg": e.bg,
                    f"iekf_{tag}_g": e.g, f"iekf_{tag}_cov": e.cov, f"iekf_{tag}_frame_q": r["frame_q"], f"iekf_{tag}_frame_t": r["frame_t"],
                    f"iekf_{tag}_num_residuals": np.array(r["num_residuals_used"])})
    # gridSampling, distortFrameBy*, transformAllImuPoint (src/utility.cpp)
    st = [dict(timestamp=r[0], quat=r[1:5], trans=r[5:8], vel=r[8:11], un_acc=r[11:14], un_gyr=r[14:17]) for r in sp["imu_states"]]
    t0 = float(sp["t0"])
    out["by_const"] = Rf.distort_frame_by_constant(sp["raw"], sp["rel"], st, t0, sp["R_il"], sp["t_il"])
    out["by_imu"] = Rf.distort_frame_by_imu(sp["raw"], sp["rel"], st, t0, sp["R_il"], sp["t_il"])
    out["by_imu_cut"] = Rf.distort_frame_by_imu(sp["raw"], sp["rel_cut"], st, t0, sp["R_il"], sp["t_il"], imu_xyz_in=np.full_like(sp["raw"], 9.0))
    out["end_frame"] = Rf.transform_all_imu_point(out["by_imu"], st[-1], sp["R_il"], sp["t_il"])
    out["grid_keep"] = Rf.grid_sampling(out["by_imu"], float(sp["grid_size"]))
    # addPointsToMap (src/lioOptimization.cpp:520-554), two batches
    m = Rf.Reference()
    out["added_a"] = np.array(m.add_points_to_map(mi["batch_a"]))
    s = m.snapshot()
    out.update(keys_a=s["keys"], counts_a=s["counts"], xyz_a=s["xyz"])
    out["added_b"] = np.array(m.add_points_to_map(mi["batch_b"]))
    s = m.snapshot()
    out.update(keys_b=s["keys"], counts_b=s["counts"], xyz_b=s["xyz"])
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("reference_outputs.npz:", {k: v.shape for k, v in out.items() if k.startswith("pass_") and k.endswith("rows")},
          os.path.getsize(os.path.join(HERE, "reference_outputs.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
