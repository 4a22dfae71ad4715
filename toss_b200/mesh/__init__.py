from .mesh_utility import create_mesh, is_outside, read_pk_file

__all__ = ["create_mesh", "is_outside", "read_pk_file"]
