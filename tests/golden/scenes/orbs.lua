-- A programmatic scene: exercises closures, loops, `require`, `:with`, arithmetic on expressions and
-- hand-written tables for objects the DSL has no helper for (directional light, bounds.sphere, clamp).
local palette = require "palette"

local ball = shape.sphere {radius = 0.6, position = vector(0, 0.6, 0)}
local objects = {}

local function add(object)
    objects[#objects + 1] = object
    return object
end

add(shape.plane {origin = vector(), normal = vector {y = 1}, material = {surface = material.diffuse {color = palette.grey}}})

local tints = {palette.warm, palette.cool, rgb(0.2, 0.9, 0.3)}
for i, tint in ipairs(tints) do
    local x = (i - 2) * 1.5
    add(ball:with {
        position = ball.position:with {x = x},
        material = {surface = mix(material.mirror {color = 1}, material.diffuse {color = tint}, fresnel(1.4))},
    })
end

add(ball:with(function(b)
    return {radius = b.radius * 0.5, position = vector(0, 2.2, -0.5), material = {surface = material.emissive {color = palette.lamp}}}
end))

-- no DSL helper exists for these: plain tables with a `type` tag
add {type = "directional_light", direction = vector(0.3, 1, 0.2), width = 0.98, color = light_source.a * 0.5}
add(shape.ray_marched {
    shape = ray_marched.mandelbulb {iterations = 6, threshold = 4, power = 8},
    bounds = {type = "sphere", position = vector(0, 1.2, 2.5), radius = 1.25},
    material = {surface = material.diffuse {color = {type = "clamp", value = palette.warm * 2, min = 0, max = 1}}},
})

local size = 48
return {
    image = {width = size * 2, height = size},
    renderer = renderer.simple {pixel_samples = 4, spectrum_samples = 4, light_samples = 2, bounces = 4, spectrum_bins = 50},
    camera = camera.perspective {fov = 45, transform = transform.look_at {from = vector(0, 1.5, -6), to = vector(0, 0.8, 0)}},
    world = {sky = light_source.d65 * 0.2, objects = objects},
}
