"""``pebble.ProcessPool`` look-alike (subset used by ``qgs/inner_products/symbolic.py:243-1690``).

``pool.map(func, iterable, timeout=...)`` returns an object whose ``.result()`` is an iterator over
results in submission order.  Work is fanned over ``multiprocessing`` workers; the per-item timeout of
the real pebble is not enforced (the reference only uses it to fall back from symbolic to numeric
integration, ``symbolic.py:1646-1663``).
"""
import multiprocessing
import os

__all__ = ["ProcessPool"]


class _MapFuture:
    def __init__(self, it):
        self._it = it

    def result(self):
        return self._it


class ProcessPool:
    def __init__(self, max_workers=None, **_ignored):
        self._n = max_workers or os.cpu_count() or 1
        self._pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def _ensure(self):
        if self._pool is None and self._n > 1 and os.environ.get("QGSB_SERIAL_POOL", "0") != "1":
            self._pool = multiprocessing.get_context("fork").Pool(self._n)
        return self._pool

    def map(self, func, iterable, timeout=None, chunksize=1):
        items = list(iterable)
        pool = self._ensure()
        if pool is None or len(items) < 2:
            return _MapFuture(iter([func(x) for x in items]))
        return _MapFuture(pool.imap(func, items, chunksize=max(1, len(items) // (8 * self._n))))

    def close(self):
        if self._pool is not None:
            self._pool.close()
            self._pool.join()
            self._pool = None

    stop = close

    def join(self):
        pass
