"""Run one of the reference's training scripts UNCHANGED over this repo's surface:

    cd /path/to/U2MKD && python -m u2mkd_b200.shims.launch [--synthetic N_TRAIN,N_VAL[,MAX_POINTS]] train_spformer.py \\
        configs/nuscenes/train/spformer.yaml --run-dir runs/x [--non-dist] [--key.subkey value ...]

What it does before handing control to the script (`runpy`, `__main__`): registers u2mkd_b200.torchsparse as `torchsparse`
(unless a `torchsparse` is already in sys.modules), puts the script's directory first on sys.path (the reference imports
`core.*`, `third_party.*`, `visualize_utils` from its checkout root), registers the stand-ins of install_reference_shims()
for the third-party packages that are not importable (torchpack, timm, torch_scatter, prettytable, nuscenes-devkit's
ConfusionMatrix, open3d's only user visualize_utils), and — with --synthetic, or when configs.dataset.root does not exist —
replaces core.builder.make_dataset by the synthetic nuScenes adapter (synthetic_nusc.py), sized by the flag.  Multi-GPU:
launch with torchrun (`torchrun --nproc-per-node N -m u2mkd_b200.shims.launch ...`); torchpack.distributed then reads
RANK / WORLD_SIZE / LOCAL_RANK and rendezvous over env:// instead of MPI."""
from __future__ import annotations

import os
import runpy
import sys


def _patch_dataset(n_train: int, n_val: int, max_points: int, force: bool, image_size=None) -> None:
    import importlib
    builder = importlib.import_module("core.builder")
    real = builder.make_dataset

    def make_dataset(dataset_name: str = None, **kwargs):
        from torchpack.utils.config import configs
        from .synthetic_nusc import SyntheticNuScenes
        root = configs.get("dataset", {}).get("root")
        if not force and root and os.path.isdir(str(root)):
            return real(dataset_name, **kwargs)
        ds = configs.dataset
        common = dict(voxel_size=ds.voxel_size, num_train=n_train, num_val=n_val,
                      multisweeps=ds.get("multisweeps", {}).get("num_sweeps", 0), num_classes=configs.data.num_classes,
                      ignored_label=configs.data.get("ignore_label", 0), seed=configs.get("train", {}).get("seed", 0) or 0,
                      max_points=max_points)
        name = dataset_name if dataset_name is not None else ds.name
        if name == "lc_semantic_nusc_tsd_full":     # LiDAR + six cameras, student / teacher inputs
            from .synthetic_nusc import SyntheticNuScenesCameras
            size = image_size or [int(x * ds.get("im_cr", 0.4)) for x in (900, 1600)]
            return SyntheticNuScenesCameras(image_size=size, im_drop=ds.get("im_drop", 0),
                                            debug=configs.get("debug", {}).get("debug_val", True), **common)
        return SyntheticNuScenes(flip_aug=ds.get("flip_aug", True), rotate_aug=ds.get("rotate_aug", True),
                                 translate_std=ds.get("translate_std", None), **common)

    builder.make_dataset = make_dataset


def run_script(script: str, argv, synthetic=None) -> None:
    """script: path of the reference script; argv: its own arguments; synthetic: None (real dataset unless its root is
    missing) or (n_train, n_val, max_points[, image_h, image_w])."""
    import u2mkd_b200
    script = os.path.abspath(script)
    root = os.path.dirname(script)
    if root not in sys.path:
        sys.path.insert(0, root)
    if "torchsparse" not in sys.modules:
        u2mkd_b200.install_as_torchsparse()
    u2mkd_b200.install_reference_shims()
    n_train, n_val, max_points = (tuple(synthetic) + (0,))[:3] if synthetic is not None else (64, 16, 0)
    image_size = tuple(synthetic[3:5]) if synthetic is not None and len(synthetic) >= 5 else None
    _patch_dataset(n_train, n_val, max_points, force=synthetic is not None, image_size=image_size)
    sys.argv = [script] + list(argv)
    runpy.run_path(script, run_name="__main__")


def main() -> None:
    args = sys.argv[1:]
    synthetic = None
    if args and args[0] == "--synthetic":
        parts = [int(v) for v in args[1].split(",")]
        synthetic = (parts[0], parts[1] if len(parts) > 1 else max(1, parts[0] // 4), parts[2] if len(parts) > 2 else 0) + tuple(parts[3:5])
        args = args[2:]
    if not args:
        raise SystemExit(__doc__)
    run_script(args[0], args[1:], synthetic)


if __name__ == "__main__":
    main()
