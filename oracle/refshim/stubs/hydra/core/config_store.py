class ConfigStore:
    _inst = None

    def __init__(self):
        self.repo = {}

    @classmethod
    def instance(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def store(self, name=None, node=None, group=None, **kw):
        self.repo[(group, name)] = node
