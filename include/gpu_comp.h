/* Device lifecycle hooks that the reference's test programs call directly
 * (/root/reference/gpu_comp.h:11-13; time_results.c:87-88,134-135;
 * compare_results.c:94,137).  The reference header also pulls in <CL/opencl.h> and
 * exposes an OpenCL context; neither exists here — this is the CUDA replacement. */
#ifndef GPU_COMP
#define GPU_COMP
#ifdef __cplusplus
extern "C" {
#endif
extern void gpu_init(void);      /* pick the device, create streams; idempotent          */
extern void gpu_cleanup(void);   /* run registered hooks, free device workspaces         */
extern void register_cleanup(void (*f)(void));
#ifdef __cplusplus
}
#endif
#endif
