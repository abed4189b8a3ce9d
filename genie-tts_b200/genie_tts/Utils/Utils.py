"""Small host utilities (reference: src/genie_tts/Utils/Utils.py:5-28)."""
import queue
from collections import OrderedDict


class LRUCacheDict(OrderedDict):
    """Dict with a capacity: reads refresh recency, inserts evict the least recently used."""

    def __init__(self, capacity: int, on_evict=None):
        super().__init__()
        self.capacity = max(1, int(capacity))
        self.on_evict = on_evict          # called with (key, value) for entries the capacity pushes out

    def __getitem__(self, key):
        val = OrderedDict.__getitem__(self, key)
        self.move_to_end(key)
        return val

    def __setitem__(self, key, value):
        OrderedDict.__setitem__(self, key, value)
        self.move_to_end(key)
        while len(self) > self.capacity:
            k, v = self.popitem(last=False)
            if self.on_evict is not None:
                self.on_evict(k, v)


def clear_queue(q: "queue.Queue") -> None:
    try:
        while True:
            q.get_nowait()
    except queue.Empty:
        pass
