"""Drop-in mirror of the reference's ``deeplens`` package for the focal-stack synthesis path.

Only the modules on the hot path exist here (psfnet, psfnet_arch, render_psf); the ray tracer,
plotting and metric helpers of the reference are out of scope.
"""
