def use(*_a, **_k):
    pass
