class _Meta:
    def get(self, name):
        raise KeyError(name)


MetadataCatalog = _Meta()
