class _Null:
    def __call__(self, *a, **k):
        return _Null()

    def __getattr__(self, name):
        return _Null()

    def __iter__(self):
        return iter((_Null(), _Null()))

    def __getitem__(self, k):
        return _Null()


def __getattr__(name):
    return _Null()
