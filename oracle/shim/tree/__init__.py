"""TEST INFRASTRUCTURE ONLY.  Minimal stand-in for dm-tree (`import tree`) as
used by the reference layer (map_structure / flatten over a bare tensor or
nested list/tuple/dict)."""


def flatten(structure):
    if isinstance(structure, (list, tuple)):
        out = []
        for s in structure:
            out.extend(flatten(s))
        return out
    if isinstance(structure, dict):
        out = []
        for k in sorted(structure):
            out.extend(flatten(structure[k]))
        return out
    return [structure]


def map_structure(func, *structures):
    s0 = structures[0]
    if isinstance(s0, (list, tuple)):
        return type(s0)(map_structure(func, *xs) for xs in zip(*structures))
    if isinstance(s0, dict):
        return {k: map_structure(func, *(s[k] for s in structures)) for k in s0}
    return func(*structures)
