class Ellipse:
    def __init__(self, *a, **k):
        pass
