"""B200-native (sm_100a) implementation of the U-Net hot path of
a-green-hand-jack/mdeical_image_segmentation: drop-in ``nn.Module`` classes, losses and metrics whose
arithmetic runs in hand-written CUDA (tcgen05/TMEM/TMA implicit-GEMM convolutions + HBM-bound fused
kernels) behind a C ABI (``include/b200unet.h``).  There is no CPU / PyTorch-eager fallback."""
from . import _lib, augment, metrics  # noqa: F401
from .graph import GraphedInference, GraphedTrainStep
from .unet3plus import UNet_3Plus, UNet_3Plus_DeepSup, UNet_3Plus_DeepSup_CGM
from .losses import (BCEDiceLoss, BCEWithLogitsLoss, CrossEntropyLoss, DiceLoss, F1Loss, IoULoss, MSSSIMLoss,
                     SegmentationLoss, compute_per_channel_dice, flatten, get_loss_criterion)
from .unet2d import (CropAndConcat, DoubleConvolution, DownSample, UNet, UpSample, unetConv2)
from .unet3d import (AbstractUNet, ChannelSELayer3D, ChannelSpatialSELayer3D, Decoder, DoubleConv, Encoder,
                     ResidualUNet2D, ResidualUNet3D, ResidualUNetSE3D, ResNetBlock, ResNetBlockSE, SingleConv,
                     SpatialSELayer3D, UNet2D, UNet3D, create_conv, create_decoders, create_encoders, get_model,
                     number_of_features_per_level)
from .predictor import LazyPredictor, StandardPredictor
from .serving import SegmentationService, create_app

try:
    from .unet2d import UNetConfig, UNetModel, UNetModelOutput
    from .unet3d import (UNet3DForMedicalSegmentation, UNet3DForMedicalSegmentationConfig,
                         UNet3DForMedicalSegmentationOutput)
except ImportError:  # transformers missing
    pass

__version__ = "0.1.0"
