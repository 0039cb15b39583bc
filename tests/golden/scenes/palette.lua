-- shared colours for the test scenes (loaded with `require "palette"`)
local function band(center, width, peak)
    return spectrum {
        format = "curve",
        points = {{center - width, 0}, {center - width / 2, peak * 0.6}, {center, peak}, {center + width / 2, peak * 0.6}, {center + width, 0}},
    }
end

return {
    warm = band(610, 60, 0.9),
    cool = band(470, 50, 0.8),
    grey = 0.5,
    lamp = light_source.d65 * 4,
}
