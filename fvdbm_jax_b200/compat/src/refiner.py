"""reference src/refiner.py (MeshRefiner: quality metrics, smoothing, edge flips, patch remeshing on top of
meshpy.triangle).  Offline mesh generation is outside this framework's scope; the name exists
so that the notebooks' ``from src.refiner import *`` line imports, and says what to do instead when it is used."""

__all__ = ["MeshRefiner"]


class MeshRefiner:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "MeshRefiner needs meshpy.triangle and is not part of fvdbm_jax_b200: refine the mesh with the reference's "
            "src/refiner.py (or any mesher) and hand the result to Mesher.import_meshpy, or use fvdbm_jax_b200.meshgen")
