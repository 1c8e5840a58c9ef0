def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)
