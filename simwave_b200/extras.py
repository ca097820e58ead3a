"""
Names simwave exports that are NOT on the forward-modelling hot path
(simwave/io/io.py, simwave/plots/plot.py).  They are out of scope for this
package and exist only so that
``from simwave_b200 import *`` offers the same names; each one tells the user
where the functionality lives.
"""


def _out_of_scope(name, origin):
    def stub(*args, **kwargs):
        raise NotImplementedError(
            "{}() is not part of the B200 forward-modelling backend; use "
            "simwave's own {}.".format(name, origin)
        )
    stub.__name__ = name
    stub.__doc__ = "Out of scope; see simwave {}.".format(origin)
    return stub


read_2D_segy = _out_of_scope("read_2D_segy", "simwave/io/io.py")
plot_wavefield = _out_of_scope("plot_wavefield", "simwave/plots/plot.py")
plot_shotrecord = _out_of_scope("plot_shotrecord", "simwave/plots/plot.py")
plot_velocity_model = _out_of_scope("plot_velocity_model",
                                    "simwave/plots/plot.py")
plot_wavelet = _out_of_scope("plot_wavelet", "simwave/plots/plot.py")
