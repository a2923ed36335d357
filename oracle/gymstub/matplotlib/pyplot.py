"""Recording stand-in for matplotlib.pyplot: every call the reference's render() makes (pogostick_v1_env.py:556-620) is
appended to CALLS as (name, args, kwargs), so that oracle/gen_render_golden.py can pin what the reference draws."""
CALLS = []


def _recorder(name):
    def call(*args, **kwargs):
        CALLS.append((name, args, kwargs))
    return call


for _name in ('figure', 'imshow', 'arrow', 'title', 'xlabel', 'ylabel', 'text', 'legend', 'tight_layout', 'pause', 'clf',
              'show', 'savefig', 'colorbar', 'grid'):
    globals()[_name] = _recorder(_name)
