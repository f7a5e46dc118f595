from .._core import (CaptionProjection, GaussianFourierProjection, ImageHintTimeEmbedding, ImageProjection,  # noqa: F401
                     ImageTimeEmbedding, PositionNet, SinusoidalPositionalEmbedding, TextImageProjection,
                     TextImageTimeEmbedding, TextTimeEmbedding, TimestepEmbedding, Timesteps, get_timestep_embedding)
