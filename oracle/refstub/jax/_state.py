"""global switches of the jax stand-in"""
import torch

x64 = False


def default_float():
    return torch.float64 if x64 else torch.float32


def default_int():
    return torch.int64 if x64 else torch.int32
