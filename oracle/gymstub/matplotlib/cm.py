"""matplotlib.cm stand-in: a colour map is the function fraction -> ('cmap', name, fraction), enough to pin the legend."""


def get_cmap(name=None, *a, **k):
    def cmap(fraction):
        return ('cmap', name, round(float(fraction), 9))
    return cmap
