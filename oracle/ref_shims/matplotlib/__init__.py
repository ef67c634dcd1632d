"""Empty stand-in so the reference's plotting imports resolve (oracle/ref_harness.py).
TEST INFRASTRUCTURE; the hot path never calls into matplotlib."""


def use(*args, **kwargs):
    return None
