"""Development helper: stage-by-stage comparison of the projective path against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.set_num_threads(1)
import pylidar_slam_b200 as b200
from pylidar_slam_b200 import synthetic as syn
from oracle import icp_oracle as orc

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 2048)
vm0 = syn.vertex_map_from_scan(syn.scan(0, H, W), H, W)
vm1 = syn.vertex_map_from_scan(syn.scan(1, H, W), H, W)
proj = b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0)
mine = b200.ProjectiveLocalMap(b200.ProjectiveLocalMapConfig(local_map_size=20), projector=proj)
theirs = orc.ProjectiveLocalMap(orc.Projector(H, W), local_map_size=20)
mine.init()
eye = np.eye(4, dtype=np.float32)
mine.update(eye[None], new_vertex_map=vm0)
theirs.update(torch.from_numpy(eye)[None], new_vertex_map=torch.from_numpy(vm0))
mv, mn = mine.model()
tv, tn = theirs.model_vmap.numpy(), theirs.model_nmap.numpy()
print("model vmap exact-equal frac", np.mean(mv == tv), "nmap exact-equal frac", np.mean(mn == tn))
bad = np.any(mv != tv, axis=1)
print("  differing pixels", bad.sum(), "of", bad.size)
q = torch.from_numpy(vm1)[0].permute(1, 2, 0).reshape(-1, 3)
q = q[q.norm(dim=-1) > 0].numpy()
res = mine.nearest_neighbor_search(q)
tq, tnn, tp = theirs.nearest_neighbor_search(torch.from_numpy(q))
print("Nc mine", res.neighbor_points.shape[1], "oracle", tq.shape[1])
if res.neighbor_points.shape[1] == tq.shape[1]:
    print("  targets equal frac", np.mean(res.new_target_points[0] == tp[0].numpy()),
          "neighbors equal frac", np.mean(res.neighbor_points[0] == tq[0].numpy()),
          "normals equal frac", np.mean(res.neighbor_normals[0] == tnn[0].numpy()))
al = b200.GaussNewtonPointToPlaneAlignment(b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)))
dT, x, loss = al.align(res.neighbor_points, res.new_target_points, res.neighbor_normals)
dTo, xo, losso = orc.align_p2plane(tq, tp, tnn, "geman_mcclure", 0.3, 1)
print("GN on own correspondences: x mine", x[0], "oracle", xo[0].numpy(), "loss", loss.sum(), float(losso.sum()))
# full ICP register on the same map state
cfg = b200.ICPFrameToModelConfig(local_map=b200.ProjectiveLocalMapConfig(local_map_size=20),
    alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
    max_num_alignments=10, data_key="vertex_map")
algo = b200.ICPFrameToModel(cfg, projector=proj, device="cuda:0"); algo.init()
algo.process_next_frame({"vertex_map": torch.from_numpy(vm0)})
params, T, losses = algo.register_new_frame(q, None)
ocfg = orc.ICPConfig(max_num_alignments=10, data_key="vertex_map", local_map="projective", local_map_size=20, scheme="geman_mcclure", sigma=0.3)
oalgo = orc.ICPFrameToModelOracle(ocfg, orc.Projector(H, W))
oalgo.process_next_frame({"vertex_map": torch.from_numpy(vm0)})
op, oT, ol = oalgo.register_new_frame(torch.from_numpy(q), torch.eye(4).unsqueeze(0))
print("losses mine  ", np.array(losses))
print("losses oracle", np.array(ol))
print("T mine\n", T[0], "\nT oracle\n", oT[0].numpy())
