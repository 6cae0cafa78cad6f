"""inert stand-in (abstract_cyclic_gen.py imports matplotlib for its plot() helper only)"""
def use(*a, **k):
    pass
