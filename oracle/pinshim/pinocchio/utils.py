"""pin.utils.zero, as the reference's motion files use it (TEST INFRASTRUCTURE ONLY)"""
import numpy as np


def zero(n):
    return np.zeros(n)
