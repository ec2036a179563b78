__sylph_shim__ = True  # stand-in package: the product must never bind to it (see structures.py)
