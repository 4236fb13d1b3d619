"""Problem definitions shared by oracle/make_golden.py (which runs the reference on them) and the parity tests.
Material data and boundary-condition style follow the reference apps' __main__ blocks
(apps/heat_transfer.py:165-186, apps/stress_equilibrium.py:191-218, apps/geomechanics.py:287-329)."""

N0 = ("neumann", 0.0)
HEAT_PROPS = {"Body": {"HeatCapacity": 1.0, "Conductivity": 1.0, "Density": 1.0, "HeatGeneration": 100.0}}
_we = {"West": ("dirichlet", 300.0), "East": ("dirichlet", 400.0), "South": N0, "North": N0}
_we3 = dict(_we, Bottom=N0, Top=N0)
BC3D = {"West": ("dirichlet", 300.0), "East": ("dirichlet", 400.0), "Bottom": ("dirichlet", 500.0),
        "Top": ("neumann", 1000.0), "South": N0, "North": N0}
_two = {"WEST": {"HeatCapacity": 1.0, "Conductivity": 1.0, "Density": 1.0, "HeatGeneration": 100.0},
        "EAST": {"HeatCapacity": 2.0, "Conductivity": 5.0, "Density": 3.0, "HeatGeneration": 10.0}}

HEAT_CASES = {
    "Square": dict(mesh="2D/Square.msh", props=HEAT_PROPS, bcs=_we, transient_steps=2),
    "10x10": dict(mesh="2D/10x10.msh", props=HEAT_PROPS, bcs=_we, transient_steps=0),
    "test_mixed": dict(mesh="2D/test.msh", props=_two,
                       bcs={"SW": N0, "W": ("dirichlet", 300.0), "NW": ("neumann", 50.0), "MW": N0, "SE": N0,
                            "E": ("dirichlet", 400.0), "NE": ("neumann", -20.0), "ME": N0}, transient_steps=2),
    "Hexas3x3x3_we": dict(mesh="3D/Hexas3x3x3.msh", props=HEAT_PROPS, bcs=_we3, transient_steps=0),
    "Hexas3x3x3_3d": dict(mesh="3D/Hexas3x3x3.msh", props=HEAT_PROPS, bcs=BC3D, transient_steps=2),
    "Hexas6x6x6_we": dict(mesh="3D/Hexas6x6x6.msh", props=HEAT_PROPS, bcs=_we3, transient_steps=0),
    "eight_sphere": dict(mesh="3D/eight_sphere.msh", props=HEAT_PROPS,
                         bcs={"Outer": ("dirichlet", 300.0), "InnerZ": ("neumann", 10.0), "InnerY": ("dirichlet", 350.0),
                              "InnerX": N0}, transient_steps=2),
    "Tetras_we": dict(mesh="3D/Tetras.msh", props=HEAT_PROPS, bcs=_we3, transient_steps=0),
    "Prisms_3d": dict(mesh="3D/Prisms.msh", props=HEAT_PROPS, bcs=BC3D, transient_steps=2),
    "Pyrams_3d": dict(mesh="3D/Pyrams.msh", props=HEAT_PROPS, bcs=BC3D, transient_steps=2),
}

STRESS_PROPS = {"Body": {"Density": 1800.0, "PoissonsRatio": 0.4, "ShearModulus": 6.0e6, "Gravity": 10.0}}
_s2 = {"u": {"West": ("dirichlet", 0.0), "East": ("dirichlet", 0.0), "South": N0, "North": N0},
       "v": {"West": N0, "East": N0, "South": ("dirichlet", 0.0), "North": ("neumann", -1e4)}}
_s3 = {"u": dict(_s2["u"], Bottom=N0, Top=N0), "v": dict(_s2["v"], Bottom=N0, Top=N0),
       "w": {"West": N0, "East": N0, "South": N0, "North": N0, "Bottom": ("dirichlet", 0.0), "Top": N0}}
STRESS_CASES = {
    "Square": dict(mesh="2D/Square.msh", props=STRESS_PROPS, bcs=_s2, gravity=False),
    "10x10_gravity": dict(mesh="2D/10x10.msh", props=STRESS_PROPS, bcs=_s2, gravity=True),
    "Hexas4x4x4": dict(mesh="3D/Hexas4x4x4.msh", props=STRESS_PROPS, bcs=_s3, gravity=False),
    "Prisms": dict(mesh="3D/Prisms.msh", props=STRESS_PROPS, bcs=_s3, gravity=True),
    "Pyrams": dict(mesh="3D/Pyrams.msh", props=STRESS_PROPS, bcs=_s3, gravity=False),
    "eight_sphere": dict(mesh="3D/eight_sphere.msh", props=STRESS_PROPS, gravity=False,
                         bcs={"u": {"Outer": N0, "InnerZ": N0, "InnerY": N0, "InnerX": ("dirichlet", 0.0)},
                              "v": {"Outer": N0, "InnerZ": N0, "InnerY": ("dirichlet", 0.0), "InnerX": N0},
                              "w": {"Outer": ("neumann", 2e3), "InnerZ": ("dirichlet", 0.0), "InnerY": N0, "InnerX": N0}}),
}

BIOT_PROPS = {"Body": {"nu": 0.2, "G": 6.0e9, "cs": 0.0, "phi": 0.19, "k": 1.9e-15, "cf": 3.0303e-10, "mu": 1000.0,
                       "rhos": 2700.0, "rhof": 1000.0}}
_b2 = {"u_x": {"West": ("dirichlet", 0.0), "East": ("dirichlet", 0.0), "South": N0, "North": N0},
       "u_y": {"West": N0, "East": N0, "South": ("dirichlet", 0.0), "North": ("neumann", 1e5)},
       "p": {"West": N0, "East": N0, "South": N0, "North": ("dirichlet", 0.0)}}
_b3 = {"u_x": dict(_b2["u_x"], Bottom=N0, Top=N0), "u_y": dict(_b2["u_y"], Bottom=N0, Top=N0),
       "u_z": {"West": N0, "East": N0, "South": N0, "North": N0, "Bottom": ("dirichlet", 0.0), "Top": N0},
       "p": dict(_b2["p"], Bottom=N0, Top=N0)}
BIOT_CASES = {
    "Square": dict(mesh="2D/Square.msh", props=BIOT_PROPS, bcs=_b2, dt=20.0, steps=2),
    "Hexas3x3x3": dict(mesh="3D/Hexas3x3x3.msh", props=BIOT_PROPS, bcs=_b3, dt=20.0, steps=2),
    "Prisms": dict(mesh="3D/Prisms.msh", props=BIOT_PROPS, bcs=_b3, dt=20.0, steps=1),
    "Pyrams": dict(mesh="3D/Pyrams.msh", props=BIOT_PROPS, bcs=_b3, dt=20.0, steps=1),
    "eight_sphere": dict(mesh="3D/eight_sphere.msh", props=BIOT_PROPS, dt=20.0, steps=1,
                         bcs={"u_x": {"Outer": N0, "InnerZ": N0, "InnerY": N0, "InnerX": ("dirichlet", 0.0)},
                              "u_y": {"Outer": N0, "InnerZ": N0, "InnerY": ("dirichlet", 0.0), "InnerX": N0},
                              "u_z": {"Outer": ("neumann", 1e5), "InnerZ": ("dirichlet", 0.0), "InnerY": N0, "InnerX": N0},
                              "p": {"Outer": ("dirichlet", 0.0), "InnerZ": N0, "InnerY": N0, "InnerX": N0}}),
}
