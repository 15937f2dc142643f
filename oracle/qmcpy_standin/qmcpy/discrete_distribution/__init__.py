"""Stand-in base class (tests only)."""


class AbstractDiscreteDistribution(object):
    def __init__(self, dimension=1, replications=None, seed=None, d_limit=None, n_limit=None):
        self.d = int(dimension)
        self.replications = 1 if replications is None else int(replications)
        self.seed = seed

    def __call__(self, n=None, n_min=None, n_max=None, return_binary=False, **kw):
        return self._gen_samples(n_min, n_max, False, return_binary, False)[0]
