from .camera_encoder import InvertedResidual, TwinLiteEncoder  # noqa: F401
from .lidar_encoder import LiDAREncoder, SpatialLiDAREncoder, create_test_point_cloud  # noqa: F401
from .fusion_module import (CameraFPNLite, CompleteSegmentationModel, ConcatenationFusion, Conv1x1,  # noqa: F401
                            DWSeparableConv, LightweightSegmentationHead, MinimalFusion,
                            SameResolutionSegmentationHead, WeightedFusion)
