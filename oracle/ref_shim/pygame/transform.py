"""pygame.transform stand-in: only what module/constructor code of the reference touches."""


def scale(surface, size):
    from . import Surface
    return Surface(size)


def rotate(surface, angle):
    raise NotImplementedError("pygame.transform.rotate is not restated (observation rendering is not run under the shim)")
