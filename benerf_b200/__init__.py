"""benerf_b200: Blackwell-native render-and-image-formation engine behind BeNeRF's call surface.

Host-side mirror of the reference modules this path touches (same names, argument meaning and
error behaviour), every arithmetic step forwarded to libbenerf_b200.so through the C ABI:

    reference                      here
    model/nerf.py     NeRF, Graph  benerf_b200.nerf
    model/optimize.py Model, Graph benerf_b200.optimize
    model/component.py holders     benerf_b200.component
    spline.py                      benerf_b200.spline
    run_nerf_helpers.py (init, eval drivers)  benerf_b200.run_nerf_helpers
    utils/event_utils.py accumulate_events_on_gpu, train.py image formation  benerf_b200.image_formation
"""
from ._lib import BnrfError, LIB_PATH  # noqa: F401
