import numpy as np


def log(x):
    return np.log(np.asarray(x))


def cumprod(x, axis=0):
    return np.cumprod(np.asarray(x), axis=axis)


def floordiv(a, b):
    return np.asarray(a) // b
