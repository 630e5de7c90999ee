"""Head-less no-op stand-in used only when matplotlib is not installed (figures are out of scope)."""


def rc(*a, **k):
    pass


def use(*a, **k):
    pass


rcParams = {}
