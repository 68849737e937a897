"""matplotlib.pyplot stand-in: every call is a no-op that returns another no-op object (the reference's tests and
clouds call visualize_* helpers; nothing on the hot path depends on what they draw)."""


class _Nothing:
    def __call__(self, *a, **k):
        return _Nothing()

    def __getattr__(self, name):
        return _Nothing()

    def __iter__(self):
        return iter((_Nothing(), _Nothing()))

    def __getitem__(self, k):
        return _Nothing()


def __getattr__(name):
    return _Nothing()
