"""Process-wide cache for expensive host-side tables (monotonic operators).  Mirrors scarlet/cache.py:1-29."""


class Cache:
    _store = {}

    @staticmethod
    def check(name, key):
        return Cache._store[name][key]

    @staticmethod
    def set(name, key, value):
        Cache._store.setdefault(name, {})[key] = value

    @staticmethod
    def clear():
        Cache._store.clear()
