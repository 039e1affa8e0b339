"""Drop-in for the reference's core/partice_system/partice_systemv2.py (ParticleSystemV2):
ParticleSystem + scene JSON fluid blocks.  The reference's rigid-body loading is commented out in
this class (partice_systemv2.py:92-121), so rigid bodies of the scene are ignored here as well."""
from core.partice_system.partice_system import ParticleSystem


class ParticleSystemV2(ParticleSystem):
    def __init__(self, res, simulation_config, device=0):
        super().__init__(res, device=device)
        self.simulation_config = simulation_config
        self.config = simulation_config['configuration']
        self.rigidBodiesConfig = simulation_config['rigidBodies']
        self.fluidBlocksConfig = simulation_config['fluidBlocks']

    def load_rigid_body(self, rigid_body):
        """partice_systemv2.py:67-85: OBJ -> boundary points at pitch = particle diameter.  add_fluid_and_rigid never
        calls it (the call is commented out in the reference, :92-121); kept for scripts that do.  The reference's
        2D class hands a 3D mesh to trimesh here; so does this one to the voxeliser (points are (n, 3))."""
        from ti_sph_b200 import mesh
        return mesh.sample_rigid_body(rigid_body, pitch=self.particle_diameter)

    def add_fluid_and_rigid(self):
        for fluid in self.fluidBlocksConfig:                                   # :124-136
            start, end = fluid['start'], fluid['end']
            self.add_cube(lower_corner=start, cube_size=[end[0] - start[0], end[1] - start[1]],
                          material=self.material_fluid, color=0x111111, density=fluid['density'],
                          velocity=fluid['velocity'])
