from admmsolver_b200.matrix import *  # noqa: F401,F403
from admmsolver_b200 import matrix as _m
globals().update({k: getattr(_m, k) for k in dir(_m) if k.startswith('_') and not k.startswith('__')})
