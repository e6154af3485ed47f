def map_structure(fn, structure):
    if isinstance(structure, (list, tuple)):
        return type(structure)(map_structure(fn, s) for s in structure)
    return fn(structure)
