"""B200-native drop-in for the reference's ``src`` package (camera + LiDAR
distillation training hot path).  Put this directory's parent on ``sys.path`` and
``from src.models.lidar_encoder import LiDAREncoder`` etc. work as in the reference.

Importing ``src.native`` loads ``libkdfusion_b200.so``; if it has not been built the
import fails loudly -- there is no CPU or eager fallback."""
