def msd2C_fun(*a, **k):
    raise NotImplementedError("bayesmsd is not available; GenericGaussianModel is out of scope")
