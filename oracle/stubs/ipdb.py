"""Test-infrastructure stub: the reference imports ipdb only for interactive debugging."""
def set_trace(*a, **k):
    pass
