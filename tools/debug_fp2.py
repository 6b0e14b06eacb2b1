import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'omni-pq_b200'))
from oracle import pn2_oracle as O
import _pn2 as K, pointnet2_modules as M
import test_gpu_fused as T
# replay the suite order up to the FP test
for args in [(1000, 20, 36), (128, 16, 128), (4097, 260, 256), (37, 4, 4), (70000, 132, 128)]:
    T.test_gemm_plain_and_stats.__wrapped__(*args, K) if hasattr(T.test_gemm_plain_and_stats, "__wrapped__") else T.test_gemm_plain_and_stats(*args, K)
T.test_gemm_bnrelu_source_and_padding(K)
T.test_gemm_gather_source_matches_query_and_group(K, O)
T.test_transposes_roundtrip(K)
T.test_sa_config1_matches_oracle(True, K, O)
T.test_sa_config1_matches_oracle(False, K, O)
T.test_sa_three_layer_xyz_grad_matches_oracle(K, O)
T.test_sa_without_features_and_given_inds(K, O)
print("replayed; now FP")
ours, oracle = T._pair(lambda: M.PointnetFPModule(mlp=[64 + 12, 48, 20]), lambda: O.OracleFPModule(mlp=[64 + 12, 48, 20]), seed=7, randomise_bn=True)
ours.train(); oracle.train()
unknown, uf = O.uniform_cloud(2, 900, 12, seed=21)
known, kf = O.uniform_cloud(2, 150, 64, seed=22)
outs = []
for rep in range(3):
    uf_d, kf_d = uf.cuda().requires_grad_(True), kf.cuda().requires_grad_(True)
    outs.append(ours(unknown.cuda(), known.cuda(), uf_d, kf_d).detach().cpu())
oo = [oracle(unknown, known, uf.clone().requires_grad_(True), kf.clone().requires_grad_(True)).detach() for _ in range(2)]
print("ours rep0 vs rep1 equal:", torch.equal(outs[0], outs[1]), " rep1 vs rep2:", torch.equal(outs[1], outs[2]))
print("oracle rep equal:", torch.equal(oo[0], oo[1]))
for i, o in enumerate(outs):
    print("ours", i, "vs oracle0 rel", T.rel(o, oo[0]))
print("threads", torch.get_num_threads())
