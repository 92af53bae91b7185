"""Hand-off between get_embedder() and the network constructor.

create_npp_net calls get_embedder once for the NeRF-style Fourier embedder and then once per proposal, immediately
before constructing NPP_Net / NPP_Net_top1 (reference models/helpers.py:87,108-132).  The embedders register their
constants here; the network constructor consumes them to build the fused encoder table."""
import os

nerf = None            # FourierEmbedder of the current session
periodic = []          # PeriodicEmbedder objects in proposal order


def mode() -> str:
    """'coords' (default): embed() passes raw (row, col) coordinates through and the kernels encode on the fly.
    'table': embed() materialises the real 22 / 462-wide encodings in the reference layout."""
    m = os.environ.get("NPP_B200_EMBED", "coords")
    if m not in ("coords", "table"):
        raise ValueError("NPP_B200_EMBED must be 'coords' or 'table'")
    return m


def begin_session(nerf_embedder):
    global nerf, periodic
    nerf = nerf_embedder
    periodic = []


def add_periodic(e) -> int:
    periodic.append(e)
    return len(periodic) - 1
