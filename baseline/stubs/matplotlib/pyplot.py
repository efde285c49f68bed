def __getattr__(name):
    raise AttributeError(f"matplotlib stub: pyplot.{name} is not available in this image")
