"""Run the reference's REAL caller files (models/backbone_module.py, models/pq_transformer.py, unchanged, from
/root/reference or its staged copy baseline/_ref) on one of three implementations of the hot path and dump
the outputs to an .npz -- test infrastructure for tests/test_gpu_real_callers.py and tests/test_ref_glue_cpu.py.

    --impl ours    our drop-in modules (omni-pq_b200/ first on sys.path) on libpn2_b200.so
    --impl ref     the reference's own pointnet2/*.py over its own CUDA kernels (oracle/_ref/pn2_ref_ext.so
                   installed as `pointnet2._ext`) + torch/cuDNN fp32 (allow_tf32=False)
    --impl oracle  the reference's own pointnet2/*.py over the CPU oracle's ext (CPU tensors)

Each implementation runs in its own process because all three provide top-level modules of the same names
(`pointnet2_modules`, `pointnet2_utils`, `pytorch_utils`).  Weights: `--save-state` writes the model's
state_dict after construction, `--load-state` loads one (strict) -- so both sides hold identical parameters
and the state_dict contract is exercised on the way.
"""
import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "omni-pq_b200")


def ref_root():
    sys.path.insert(0, ROOT)
    from tools import stage_reference
    r = stage_reference.root()
    if r is None:
        raise SystemExit("reference callers not available (neither /root/reference nor baseline/_ref)")
    return r


def install(impl, ref):
    """Arrange sys.path / sys.modules so that `import pointnet2_modules` resolves to the chosen implementation."""
    import torch  # noqa: F401
    if impl == "ours":
        sys.path.insert(0, PKG)
    else:
        if impl == "ref":
            from oracle import build_ref
            ext = build_ref.load()
            assert ext is not None, "oracle/_ref/pn2_ref_ext.so is not built"
        else:
            from oracle import pn2_oracle
            ext = pn2_oracle.ext
        pkg = types.ModuleType("pointnet2")
        pkg.__path__ = []
        pkg._ext = ext
        sys.modules["pointnet2"] = pkg
        sys.modules["pointnet2._ext"] = ext
        sys.path.insert(0, os.path.join(ref, "pointnet2"))
    # the reference's model files append their own directories; `models` first so that `utils.pointnet_util`
    # resolves to models/utils like it does when train.py runs from the repository root
    sys.path.insert(1, os.path.join(ref, "models"))


def clouds(batch, points, seed0=1234):
    import torch
    from tools.synth_clouds import scannet_like_cloud
    return torch.stack([scannet_like_cloud(points, seed=seed0 + i) for i in range(batch)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "ref", "oracle"], required=True)
    ap.add_argument("--case", choices=["backbone_train", "pq_eval"], required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--points", type=int, default=40000)
    ap.add_argument("--save-state")
    ap.add_argument("--load-state")
    args = ap.parse_args()

    ref = ref_root()
    import numpy as np
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    install(args.impl, ref)
    dev = torch.device("cpu" if args.impl == "oracle" else "cuda")
    cloud = clouds(args.batch, args.points).to(dev)
    out = {}
    torch.manual_seed(0)
    if args.case == "backbone_train":
        from backbone_module import Pointnet2Backbone  # the reference's file
        import pointnet2_modules
        out["modules_file"] = np.array(os.path.abspath(pointnet2_modules.__file__))
        net = Pointnet2Backbone(input_feature_dim=3)
        if args.save_state:
            torch.save(net.state_dict(), args.save_state)
        if args.load_state:
            net.load_state_dict(torch.load(args.load_state), strict=True)
        net = net.to(dev).train()
        ep = net(cloud)
        cot = torch.randn(ep["fp2_features"].shape, generator=torch.Generator().manual_seed(1)).to(dev)
        (ep["fp2_features"] * cot).sum().backward()
        for k in ("sa1_inds", "sa2_inds", "fp2_inds", "sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "sa1_features",
                  "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
            out[k] = ep[k].detach().cpu().numpy()
        for n, p in net.named_parameters():
            out["grad." + n] = p.grad.detach().cpu().numpy()
        for n, b in net.named_buffers():
            out["buf." + n] = b.detach().cpu().numpy()
    else:
        sys.path.insert(2, os.path.join(ref, "scannet"))
        from pq_transformer import PQ_Transformer  # the reference's file
        import pointnet2_modules
        out["modules_file"] = np.array(os.path.abspath(pointnet2_modules.__file__))
        means = np.load(os.path.join(ref, "scannet", "meta_data", "scannet_means.npz"))["arr_0"]
        net = PQ_Transformer(input_feature_dim=3, num_class=18, num_proposal=256, num_quad_proposal=256,
                             num_heading_bin=1, num_size_cluster=18, mean_size_arr=means, sampling="vote")
        if args.save_state:
            torch.save(net.state_dict(), args.save_state)
        if args.load_state:
            net.load_state_dict(torch.load(args.load_state), strict=True)
        net = net.to(dev).eval()
        with torch.no_grad():
            ep = net({"point_clouds": cloud})
        for k in ("sa1_inds", "sa2_inds", "fp2_inds", "fp2_features", "aggregated_sample_xyz", "vote_xyz",
                  "aggregated_vote_xyz", "cluster_feature", "last_quad_scores", "last_quad_center",
                  "last_objectness_scores", "last_center", "last_sem_cls_scores"):
            out[k] = ep[k].detach().cpu().numpy()
    if dev.type == "cuda":
        torch.cuda.synchronize()
    np.savez(args.out, **out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
