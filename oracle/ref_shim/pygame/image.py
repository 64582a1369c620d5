"""pygame.image stand-in: the car sprites (car.png, car2.png) are used by human rendering only."""


def load(path):
    from . import Surface
    return Surface((30, 52))
