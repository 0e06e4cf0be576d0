"""Box space: gym's when gym (or gymnasium) is importable, else a small stand-in with the members the controller and
the driver loop read (low, high, shape, dtype, sample) -- gym is not a dependency of the accelerated path."""
import numpy as np

try:                                    # pragma: no cover - depends on the installation
    from gym import Env
    from gym.spaces import Box
except Exception:                       # noqa: BLE001
    try:                                # pragma: no cover
        from gymnasium import Env
        from gymnasium.spaces import Box
    except Exception:                   # noqa: BLE001
        class Env:                      # minimal base class with the context-manager hooks run_env calls
            metadata = {}

            def close(self):
                pass

            def __enter__(self):
                return self

            def __exit__(self, *args):
                self.close()
                return False

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low = np.asarray(low, dtype=dtype)
                self.high = np.asarray(high, dtype=dtype)
                self.shape = tuple(shape) if shape is not None else self.low.shape
                self.dtype = np.dtype(dtype)

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))
