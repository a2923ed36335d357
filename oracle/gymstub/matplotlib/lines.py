"""matplotlib.lines stand-in: Line2D keeps its keyword arguments (the legend entries of the reference's render())."""


class Line2D(object):
    def __init__(self, *a, **k):
        self.args, self.kw = a, k
