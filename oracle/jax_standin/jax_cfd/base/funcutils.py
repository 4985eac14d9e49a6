"""jax_cfd.base.funcutils.repeated as the notebooks use it (Flapping_Demo.ipynb cell 9): f applied `steps` times."""


def repeated(f, steps):
    def g(x):
        for _ in range(steps):
            x = f(x)
        return x
    return g
