-- The Cornell box of BASELINE config C1 written as a project file; the mesh comes from the committed fixture.
local spectra = require "box_spectra"

local white = {surface = material.diffuse {color = spectra.white}}
return {
    image = {width = 64, height = 64, white = blackbody(4000)},
    renderer = renderer.simple {pixel_samples = 4, spectrum_samples = 10, spectrum_bins = 50, tile_size = 32, light_samples = 1, bounces = 4},
    camera = camera.perspective {
        fov = 37.7,
        transform = transform.look_at {from = vector(-2.78, -8, 2.73), to = vector(-2.78, 0, 2.73), up = vector {z = 1}},
    },
    world = {
        objects = {
            shape.mesh {
                file = "../meshes/box.npz",
                materials = {
                    light = {surface = material.emissive {color = spectra.lamp * 3} + material.diffuse {color = 0.78}},
                    left = {surface = material.diffuse {color = spectra.red}},
                    right = {surface = material.diffuse {color = spectra.green}},
                    tall = white, short = white, back = white, ceiling = white, floor = white,
                },
            },
        },
    },
}
