"""B200-native inference hot path of the Neural_network_image_compression tf2_0 codec.

Public surface (mirrors /root/reference/tf2_0/src): Encoder, Decoder, ProClass, plus rate().
All arithmetic runs in libnnic.so (hand-written sm_100a CUDA); there is no CPU fallback.
"""
from . import container, dist, weights
from ._lib import Handle, NnicError, colour_constants, load_library
from .decoder import Decoder
from .encoder import Encoder
from .container import get_bpp, pack_latent, read_dataset, save_img, unpack_latent
from .graph import GraphCodec
from .rate import Rate, entropy_from_counts, rate, rate_channels
from .training import Entropynet, entropynet_glorot, noisy_quantise, ssim
from .utils import ProClass

__all__ = ["Encoder", "Decoder", "ProClass", "Handle", "NnicError", "Entropynet", "entropynet_glorot", "noisy_quantise", "ssim", "rate", "rate_channels", "Rate", "entropy_from_counts", "GraphCodec",
           "weights", "dist", "container", "colour_constants", "load_library", "pack_latent", "unpack_latent", "read_dataset",
           "save_img", "get_bpp"]
