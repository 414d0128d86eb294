"""Pickle-compatible stand-ins for the reference's observation normaliser objects (uhc/khrylib/utils/zfilter.py:7-73).

A reference checkpoint stores `running_state` as a pickled `uhc.khrylib.utils.zfilter.ZFilter` whose `rs` is a `RunningStat` with
the attributes `_n`, `_M`, `_S` (count, mean, sum of squared deviations); unpickling therefore needs classes of these names at this
import path, and a checkpoint written here must unpickle into the reference's classes.  Only the attribute layout is shared: the
statistics themselves live on the device (uhc_b200.nn.ZFilter); these host objects are the wire format plus a numpy fallback for
single observations (eval_seq-style loops).
"""
import numpy as np


class RunningStat:
    def __init__(self, shape):
        self._n, self._M, self._S = 0, np.zeros(shape), np.zeros(shape)

    # Chan / Welford merge of a batch [k, *shape] (k = 1 reproduces the reference's per-sample push)
    def push_batch(self, xs):
        xs = np.asarray(xs, dtype=np.float64).reshape((-1,) + self._M.shape)
        k = len(xs)
        if k == 0:
            return
        mb = xs.mean(0)
        sb = ((xs - mb) ** 2).sum(0)
        tot = self._n + k
        delta = mb - self._M
        if self._n == 0:
            self._M, self._S = mb.copy(), sb
        else:
            self._S = self._S + sb + delta * delta * (self._n * k / tot)
            self._M = self._M + delta * (k / tot)
        self._n = tot

    def push(self, x):
        self.push_batch(np.asarray(x)[None])

    n = property(lambda self: self._n)
    mean = property(lambda self: self._M)
    shape = property(lambda self: self._M.shape)

    @property
    def var(self):
        return self._S / (self._n - 1) if self._n > 1 else self._M ** 2

    @property
    def std(self):
        return np.sqrt(self.var)


class ZFilter:
    def __init__(self, shape, demean=True, destd=True, clip=10.0):
        self.demean, self.destd, self.clip = demean, destd, clip
        self.rs = RunningStat(shape)

    def __call__(self, x, update=True):
        x = np.asarray(x, dtype=np.float64)
        if update:
            self.rs.push_batch(x.reshape((-1,) + self.rs.shape))
        y = x - self.rs.mean if self.demean else x
        if self.destd:
            y = y / (self.rs.std + 1e-8)
        return np.clip(y, -self.clip, self.clip) if self.clip else y

    @classmethod
    def from_stats(cls, n, mean, S, clip=5.0):
        """host object for a checkpoint from the device statistics (n, mean[D], S[D])"""
        z = cls(np.shape(mean), clip=clip)
        z.rs._n, z.rs._M, z.rs._S = (int(n) if float(n).is_integer() else float(n)), np.array(mean, dtype=np.float64), np.array(S, dtype=np.float64)
        return z
