"""Empty matplotlib stand-in (absent from this image; plotting is out of scope).  Test infrastructure."""


def rc(*a, **k):
    pass


def use(*a, **k):
    pass


rcParams = {}
