"""matplotlib.animation stand-in: FuncAnimation draws nothing and saves nothing (see pyplot.py)."""


class FuncAnimation:
    def __init__(self, fig, func, frames=None, **kwargs):
        self.frames = frames
        if frames:
            func(0)                      # one frame is drawn so that the callback's own code runs

    def save(self, filename, **kwargs):
        pass
