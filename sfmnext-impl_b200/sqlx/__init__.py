"""sqlx -- B200-native hot path of SQLdepth self-supervised training (host side).

Python mirrors of the reference's nn.Module / Trainer signatures on top of libsqlx.so
(hand-written sm_100a CUDA behind the C ABI in include/sqlx.h).  No CPU / PyTorch fallback.
"""
from ._lib import SqlxError, lib, LIB_PATH, exported_symbols  # noqa: F401
from .photometric import (photometric_losses, reprojection_loss, depth_stats, pose_matrix,  # noqa: F401
                          smooth_loss_normalised, warp, pack_rgba, indoor_losses)
from .layers import (SSIM, BackprojectDepth, Project3D, get_smooth_loss, SILogLoss,  # noqa: F401
                     batch_post_process_disparity, predict_disparity, median_scale, median_scale_ratios, finetune_loss,
                     transformation_from_parameters, inverse_rotation_warp, euler2mat)
from .trainer import FusedLossMixin, IndoorFusedLossMixin  # noqa: F401
from .sql import (FullQueryLayer, Depth_Decoder_QueryTr, Lite_Depth_Decoder_QueryTr, sql_tail,  # noqa: F401
                  bin_centers, convert_depth_decoder, fuse_depth_decoder)
from . import dist  # noqa: F401,E402
