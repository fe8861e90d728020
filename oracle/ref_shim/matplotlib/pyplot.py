def figure(*a, **k):
    return None
