make_axes_locatable = None
