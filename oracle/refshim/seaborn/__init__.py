def set(*a, **k):  # noqa: A001
    pass


def despine(*a, **k):
    pass
