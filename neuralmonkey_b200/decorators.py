"""`@tensor`: per-batch lazily evaluated, cached attribute.

The reference's decorator (neuralmonkey/decorators.py:9-27) caches a graph node for
the lifetime of the object; here the value is a device tensor computed from the
batch currently fed, so the cache lives until the next `feed_dict` call.
"""
from functools import wraps


def tensor(func):
    name = func.__name__

    @wraps(func)
    def decorate(self):
        cache = self.__dict__.setdefault("_batch_cache", {})
        if name not in cache:
            cache[name] = func(self)
        return cache[name]

    return property(decorate)
