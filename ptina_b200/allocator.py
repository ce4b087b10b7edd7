"""Host-side bookkeeping for the texture arena (reference: ptina/allocator.py:6-53).

`MemoryAllocator` hands out texel ranges first-fit from a free list; `IdAllocator` hands out image ids from a
water mark.  Same error behaviour: RuntimeError('Out of memory!') / RuntimeError('Out of ID!').
"""


class MemoryAllocator:
    def __init__(self, size):
        self.size = size
        self.reset()

    def reset(self):
        self.free_chunk = [(0, self.size)]
        self.used_chunk = []

    def malloc(self, size):
        for k, (start, length) in enumerate(self.free_chunk):
            if length < size:
                continue
            if length == size:
                self.free_chunk.pop(k)
            else:
                self.free_chunk[k] = (start + size, length - size)
            self.used_chunk.append((start, size))
            return start
        raise RuntimeError('Out of memory!')

    def free(self, base):
        for k, (start, length) in enumerate(self.used_chunk):
            if start == base:
                self.used_chunk.pop(k)
                self.free_chunk.insert(k, (start, length))
                return
        raise RuntimeError(f'Invalid pointer: {base!r}')


class IdAllocator:
    def __init__(self, count):
        self.count = count
        self.reset()

    def reset(self):
        self.water = 0

    def malloc(self):
        if self.water >= self.count:
            raise RuntimeError('Out of ID!')
        self.water += 1
        return self.water - 1

    def free(self, id):
        pass
