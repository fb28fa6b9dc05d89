import torch

from .numpy import _f


def softmax(x, axis=-1):
    return torch.softmax(_f(x), dim=axis)


def log_softmax(x, axis=-1):
    return torch.log_softmax(_f(x), dim=axis)


def sigmoid(x):
    return torch.sigmoid(_f(x))


def relu(x):
    return torch.relu(_f(x))


def softplus(x):
    return torch.nn.functional.softplus(_f(x))
