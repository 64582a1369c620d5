from . import contactListener, fixtureDef, polygonShape, revoluteJointDef  # noqa: F401
