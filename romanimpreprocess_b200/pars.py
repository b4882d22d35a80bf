"""Detector geometry and photometric constants used on the hot path.

Same names and values as the reference's ``pars.py`` (reference:
src/romanimpreprocess/pars.py:8-21) so caller code that does ``pars.nside`` keeps working.
"""

nside = 4096
nborder = 4
nchannel = 32

nside_active = nside - 2 * nborder
channelwidth = nside // nchannel
nside_augmented = nside + channelwidth

Omega_ideal = 2.8440360952308436e-13  # (0.11 arcsec)^2 in sr
h_Planck = 6.62607015e-24  # noqa: N816
g_ideal = 1.458  # e/DN zero point of the flattened DN scale
