import os


class _PathManager:
    @staticmethod
    def exists(path):
        return bool(path) and os.path.exists(path)

    @staticmethod
    def get_local_path(path):
        return path


PathManager = _PathManager()
