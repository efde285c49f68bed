class CachedDict:
    """dict that fills missing keys by calling generator_func (tuple keys are splatted)."""

    def __init__(self, generator_func=None):
        self.generator_func = generator_func
        self._cache = {}

    def __getitem__(self, key):
        if key not in self._cache:
            if isinstance(key, tuple):
                self._cache[key] = self.generator_func(*key)
            else:
                self._cache[key] = self.generator_func(key)
        return self._cache[key]

    def __setitem__(self, key, value):
        self._cache[key] = value

    def __contains__(self, key):
        return key in self._cache

    def __iter__(self):
        return iter(self._cache)

    def __len__(self):
        return len(self._cache)

    def keys(self):
        return self._cache.keys()

    def values(self):
        return self._cache.values()

    def items(self):
        return self._cache.items()
