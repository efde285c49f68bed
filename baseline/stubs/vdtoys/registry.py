class Registry:
    """Named registry; instances created with the same name share their entries."""

    _stores = {}

    def __init__(self, name):
        self.name = name
        self._store = Registry._stores.setdefault(name, {})

    def register(self, name=None):
        def deco(obj):
            self._store[name or obj.__name__] = obj
            return obj

        return deco

    def __getitem__(self, key):
        return self._store[key]

    def __contains__(self, key):
        return key in self._store

    def keys(self):
        return self._store.keys()
