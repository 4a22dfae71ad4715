from .mesh_utility import (check_mesh, create_mesh, is_outside, mesh_from_arrays, read_mesh_file, read_npz_file,
                           read_obj_file, read_pk_file)

__all__ = ["create_mesh", "is_outside", "read_pk_file", "read_obj_file", "read_npz_file", "read_mesh_file",
           "mesh_from_arrays", "check_mesh"]
