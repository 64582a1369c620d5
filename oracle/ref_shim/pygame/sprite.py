class Sprite(object):
    def __init__(self, *groups):
        pass
