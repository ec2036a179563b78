import os


class _PathManager:
    """Local-filesystem stand-in for iopath's PathManager: the calls the reference's hot-path and evaluation-loop code makes."""

    @staticmethod
    def exists(path):
        return bool(path) and os.path.exists(path)

    @staticmethod
    def isfile(path):
        return bool(path) and os.path.isfile(path)

    @staticmethod
    def mkdirs(path):
        os.makedirs(path, exist_ok=True)

    @staticmethod
    def rm(path):
        os.remove(path)

    @staticmethod
    def open(path, mode="r"):
        return open(path, mode)

    @staticmethod
    def get_local_path(path):
        return path


PathManager = _PathManager()
