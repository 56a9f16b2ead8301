from .epipolar_sampler import EpipolarSampler, EpipolarSampling
from .epipolar_transformer import EpipolarTransformer, EpipolarTransformerCfg, ImageSelfAttentionWrapper
from .image_self_attention import ImageSelfAttention, ImageSelfAttentionCfg
from .positional_encoding import PositionalEncoding
from .transformer import Attention, FeedForward, PreNorm, Transformer
