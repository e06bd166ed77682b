"""Empty stub: the reference imports matplotlib.pyplot only for plotting."""
