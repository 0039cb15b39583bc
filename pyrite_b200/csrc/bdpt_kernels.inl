// included by kernels.cu inside namespace pyr
void launch_wave_bidirectional(const SceneView&, const WaveArgs&, cudaStream_t) {}
