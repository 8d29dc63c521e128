"""Activation functions used by the flip-flop models (names as in
taiyaki/activation.py: `tanh` :123, `swish` :103, `sigmoid` :127, `linear` :14)."""
import torch


def linear(x):
    return x


def relu(x):
    return torch.relu(x)


def tanh(x):
    return torch.tanh(x)


def sigmoid(x):
    return torch.sigmoid(x)


def swish(x):
    """x * sigmoid(x) (taiyaki/activation.py:103-120), as one fused kernel"""
    return torch.nn.functional.silu(x)
