def ion():
    pass


def __getattr__(name):
    raise RuntimeError(f"matplotlib.pyplot.{name}: plotting is not part of the golden generation")
