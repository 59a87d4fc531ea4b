class _PM:
    def open(self, *a, **k):
        return open(*a, **k)
g_pathmgr = _PM()
