"""The reference's own smoke scripts, mirrored as pytest against the drop-in modules:
``test_camera_encoder.py:5-76``, ``test_lidar_encoder.py:262-313`` (live part) plus the assertions its authors left
commented out (``:1-259``), and ``test_fusion_module.py:12-66``.  Same constructor calls, same shapes, same checks --
only the device is CUDA."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def test_camera_encoder_smoke():
    """test_camera_encoder.py:5-76: 363,520 parameters; single-scale and multiscale shapes for 256^2, 512^2 and a
    batch of four 128^2 images."""
    from src.models.camera_encoder import TwinLiteEncoder
    enc = TwinLiteEncoder(in_channels=3, base_channels=32).cuda().eval()
    assert enc.count_parameters() == sum(p.numel() for p in enc.parameters() if p.requires_grad) == 363_520
    with torch.no_grad():
        for shape, want in (((2, 3, 256, 256), (2, 128, 32, 32)), ((2, 3, 512, 512), (2, 128, 64, 64)),
                            ((4, 3, 128, 128), (4, 128, 16, 16))):
            assert tuple(enc(torch.randn(*shape, device="cuda")).shape) == want
    ms = TwinLiteEncoder(in_channels=3, base_channels=32, return_multiscale=True).cuda().eval()
    with torch.no_grad():
        feats = ms(torch.randn(2, 3, 256, 256, device="cuda"))
    assert {k: tuple(v.shape) for k, v in feats.items()} == {"stage2": (2, 64, 64, 64), "stage3": (2, 64, 64, 64),
                                                             "stage4": (2, 128, 32, 32), "stage5": (2, 128, 32, 32)}
    assert ms.get_feature_info() == {"stage2": 64, "stage3": 64, "stage4": 128, "stage5": 128}


@pytest.mark.parametrize("mode,want", [("x4", (2, 3, 256, 256)), ("same", (2, 3, 64, 64))])
def test_full_model_smoke_both_heads(mode, want):
    """test_lidar_encoder.py:262-313: the complete 3-class model in ``x4`` and ``same`` output modes."""
    from src.models.camera_encoder import TwinLiteEncoder
    from src.models.fusion_module import CompleteSegmentationModel
    from src.models.lidar_encoder import LiDAREncoder, create_test_point_cloud
    model = CompleteSegmentationModel(camera_encoder=TwinLiteEncoder(return_multiscale=True),
                                      lidar_encoder=LiDAREncoder(encoder_type="spatial", grid_size=(64, 64)),
                                      num_classes=3, fusion_type="concat", fusion_out_channels=256,
                                      camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128,
                                      output_mode=mode).cuda().eval()
    images = torch.randn(2, 3, 256, 256, device="cuda")
    points = create_test_point_cloud(2, 5000, device="cuda")
    with torch.no_grad():
        logits = model(images, points)
    assert tuple(logits.shape) == want and torch.isfinite(logits).all()
    summary = model.get_architecture_summary()
    assert summary["fusion_type"] == "concat" and summary["output_mode"] == mode


def test_lidar_encoder_revived_assertions():
    """The checks the reference's authors wrote and commented out (test_lidar_encoder.py:42-44, 90-95, 122-123,
    137-142, 227-233): output shape, vectorised-vs-iterative difference, wrapper shape, PointPillars falling back to
    the spatial encoder, points outside the range giving an all-zero map."""
    from src.models.lidar_encoder import LiDAREncoder, SpatialLiDAREncoder, create_test_point_cloud
    for grid in ((128, 128), (64, 64)):
        enc = SpatialLiDAREncoder(feature_dim=128, grid_size=grid).cuda().eval()
        pts = create_test_point_cloud(2, 5000, device="cuda")
        with torch.no_grad():
            out_v = enc(pts)
            enc.use_vectorized = False
            out_i = enc(pts)
        assert tuple(out_v.shape) == (2, 128, *grid)
        assert (out_v - out_i).abs().mean().item() < 1.0          # :90-95 (ours agree bit for bit)
        assert torch.equal(out_v, out_i)
    wrapper = LiDAREncoder(encoder_type="spatial", grid_size=(64, 64)).cuda().eval()
    with torch.no_grad():
        assert tuple(wrapper(create_test_point_cloud(2, 1000, device="cuda")).shape) == (2, *wrapper.get_output_shape())
    fallback = LiDAREncoder(encoder_type="pointpillars", grid_size=(64, 64))
    assert fallback.encoder_type == "spatial" and isinstance(fallback.encoder, SpatialLiDAREncoder)
    with pytest.raises(ValueError, match="Unknown encoder type"):
        LiDAREncoder(encoder_type="voxelnet")
    far = torch.full((2, 1000, 4), 1000.0, device="cuda")
    with torch.no_grad():
        assert wrapper(far).max().item() == 0.0                   # :227-233


def test_fusion_module_smoke_forward_backward():
    """test_fusion_module.py:12-66: concat / 3-class forward with intermediates, CE backward, and "some gradient is
    non-zero in the head and in the fusion block" (:61-64).  The script's stale logits assertion (B,3,256,256) is the
    reference's own bug (default ``output_mode="same"`` gives 64x64); the real shape is checked."""
    from src.models.camera_encoder import TwinLiteEncoder
    from src.models.fusion_module import CompleteSegmentationModel
    from src.models.lidar_encoder import LiDAREncoder, create_test_point_cloud
    B = 2
    model = CompleteSegmentationModel(camera_encoder=TwinLiteEncoder(return_multiscale=True),
                                      lidar_encoder=LiDAREncoder(encoder_type="spatial", grid_size=(64, 64)),
                                      num_classes=3, fusion_type="concat", fusion_out_channels=256,
                                      camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128).cuda().train()
    images = torch.randn(B, 3, 256, 256, device="cuda")
    points = create_test_point_cloud(B, 5000, device="cuda")
    logits, mid = model(images, points, return_intermediates=True)
    assert tuple(mid["camera_feat"].shape) == (B, 128, 64, 64) and tuple(mid["lidar_feat"].shape) == (B, 128, 64, 64)
    assert tuple(mid["pre_fusion"].shape) == (B, 256, 64, 64) and tuple(mid["post_fusion"].shape) == (B, 256, 64, 64)
    assert tuple(logits.shape) == (B, 3, 64, 64)
    target = torch.randint(0, 3, (B, 64, 64), device="cuda")
    loss = nn.CrossEntropyLoss()(logits, target)
    loss.backward()
    assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in model.head.parameters())
    assert any(p.grad is not None and p.grad.abs().sum() > 0 for p in model.fusion.parameters())
    for ft in ("minimal", "weighted"):
        m = CompleteSegmentationModel(TwinLiteEncoder(return_multiscale=True), LiDAREncoder("spatial", grid_size=(64, 64)),
                                      num_classes=2, fusion_type=ft, fusion_out_channels=128,
                                      camera_fpn_stages=["stage3", "stage4", "stage5"]).cuda().eval()
        with torch.no_grad():
            assert tuple(m(images, points).shape) == (B, 2, 64, 64)
