"""Small host utilities (reference: src/genie_tts/Utils/Utils.py:5-28)."""
import queue
from collections import OrderedDict


class LRUCacheDict(OrderedDict):
    """Dict with a capacity: reads refresh recency, inserts evict the least recently used."""

    def __init__(self, capacity: int):
        super().__init__()
        self.capacity = max(1, int(capacity))

    def __getitem__(self, key):
        val = OrderedDict.__getitem__(self, key)
        self.move_to_end(key)
        return val

    def __setitem__(self, key, value):
        OrderedDict.__setitem__(self, key, value)
        self.move_to_end(key)
        while len(self) > self.capacity:
            self.popitem(last=False)


def clear_queue(q: "queue.Queue") -> None:
    try:
        while True:
            q.get_nowait()
    except queue.Empty:
        pass
