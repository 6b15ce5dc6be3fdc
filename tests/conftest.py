"""pytest configuration: markers, import paths, shared helpers."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200")
REFERENCE = "/root/reference"

for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    have_gpu = torch.cuda.is_available()
    have_ref = os.path.isdir(os.path.join(REFERENCE, "src"))
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not mounted"))


def load_reference_module(name: str):
    """Import one file of the reference under a private module name, so it can
    coexist with the product's own ``src`` package.  name e.g. 'models/lidar_encoder'."""
    path = os.path.join(REFERENCE, "src", name + ".py")
    modname = "_reference_" + name.replace("/", "_")
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def build_reference_model(fusion_type="weighted", num_classes=2, fusion_out_channels=None,
                          output_mode="same", grid_size=(64, 64)):
    """The reference's own model, wired exactly as train_with_fusion_ablation.py:27-39."""
    cam = load_reference_module("models/camera_encoder")
    lid = load_reference_module("models/lidar_encoder")
    fus = load_reference_module("models/fusion_module")
    if fusion_out_channels is None:
        fusion_out_channels = 256 if fusion_type == "concat" else 128
    return fus.CompleteSegmentationModel(
        camera_encoder=cam.TwinLiteEncoder(return_multiscale=True),
        lidar_encoder=lid.LiDAREncoder(encoder_type="spatial", grid_size=grid_size, use_vectorized=True),
        num_classes=num_classes, fusion_type=fusion_type, fusion_out_channels=fusion_out_channels,
        camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128,
        output_mode=output_mode)
