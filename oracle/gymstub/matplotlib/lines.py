class Line2D(object):
    def __init__(self, *a, **k):
        pass
