def np_random(seed=None):
    import numpy as np
    return np.random.RandomState(seed), seed
