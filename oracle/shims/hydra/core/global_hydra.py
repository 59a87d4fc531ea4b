class _GH:
    def is_initialized(self):
        return True
    def clear(self):
        pass
class GlobalHydra:
    @staticmethod
    def instance():
        return _GH()
