def MSDfun(f):
    return f


def imaging(**kw):
    def deco(f):
        return f
    return deco
