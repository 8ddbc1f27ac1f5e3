/*
 * cavity.c — the C ABI of liblife_b200 from a plain C program: a lid-driven cavity (BGK, omega = 1, lid 0.1 lattice units)
 * on one B200, with the fluid files written from the device while the time loop goes on.
 *
 *   gcc -O2 -Iinclude examples/cavity.c -Llife_b200/lib -llife_b200 -Wl,-rpath,$PWD/life_b200/lib -lm -o cavity
 *   ./cavity 4096 1000 out_dir
 *
 * What a host has to do (include/life_b200.h): describe the case in a life_config, hand over the initial state in the
 * reference's layout (f[(i*Ny + j)*9 + v], x-major), call life_step once per time step, and ask for files / scalars when it
 * wants them.  There is no CPU fallback: without a B200 life_create fails and the program stops.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "life_b200.h"

#define CK(call)                                                                        \
	do {                                                                                \
		int rc_ = (call);                                                               \
		if (rc_ != LIFE_OK) {                                                           \
			fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, life_last_error(ctx));  \
			return 99;                                                                  \
		}                                                                               \
	} while (0)

int main(int argc, char **argv) {
	const int64_t N = argc > 1 ? atoll(argv[1]) : 1024;
	const int steps = argc > 2 ? atoi(argv[2]) : 1000;
	const char *out = argc > 3 ? argv[3] : ".";
	life_ctx *ctx = NULL;

	life_config c;
	memset(&c, 0, sizeof c);
	c.abi_version = LIFE_ABI_VERSION;
	c.collision = LIFE_BGK;
	c.Nx = N; c.Ny = N;
	c.omega = 1.0;
	c.wall_left = c.wall_right = c.wall_bottom = LIFE_WALL;
	c.wall_top = LIFE_VELOCITY;
	c.inlet_ramp = -1.0; c.womersley = -1.0;
	/* the reference's scalings for height 1 m, lid 1 m/s = 0.1 lattice units (src/Grid.cpp:1257-1260) */
	c.Dx = 1.0 / (double)(N - 1);
	c.Dt = 0.1 * c.Dx;
	c.Dm = c.Dx * c.Dx * c.Dx;
	c.Drho = 1.0;
	c.device = -1;
	c.nranks = 1;
	if (life_create(&c, &ctx) != LIFE_OK) {
		fprintf(stderr, "life_create: %s\n", life_last_error(NULL));
		return 99;
	}

	/* rho = 1, u = 0: f = w (initialiseGrid, src/Grid.cpp:999-1058), streamed in column ranges so the host never holds the lattice */
	const double w[9] = {4.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 9, 1.0 / 36, 1.0 / 36, 1.0 / 36, 1.0 / 36};
	const int64_t chunk = N < 64 ? N : 64;
	double *f = malloc(sizeof(double) * (size_t)(chunk * N * 9)), *u_in = calloc((size_t)(2 * N), sizeof(double));
	for (int64_t k = 0; k < chunk * N; k++) memcpy(f + 9 * k, w, sizeof w);
	for (int64_t j = 0; j < N; j++) u_in[2 * j] = 0.1;          /* lid velocity in lattice units (GridClass::u_in) */
	CK(life_upload_begin(ctx, u_in, NULL));
	for (int64_t i0 = 0; i0 < N; i0 += chunk)
		CK(life_upload_columns(ctx, i0, i0 + chunk <= N ? chunk : N - i0, f, NULL, NULL, NULL, NULL));
	CK(life_upload_end(ctx));
	free(f); free(u_in);

	char path[1024];
	for (int t = 1; t <= steps; t++) {
		CK(life_step(ctx, t));
		if (t % (steps / 10 > 0 ? steps / 10 : 1) == 0) {         /* writeInfo: one scan on the device, three scalars back */
			double vmax; int32_t nan; int64_t ni, nj;
			CK(life_max_speed(ctx, &vmax, &nan, &ni, &nj));
			printf("t = %d  max |u| = %.6f%s\n", t, vmax, nan ? "  (NaN!)" : "");
			if (nan) return 99;
		}
		if (t % (steps / 2 > 0 ? steps / 2 : 1) == 0) {           /* writeVTK: returns at once, the worker writes while we step */
			snprintf(path, sizeof path, "%s/Fluid.%d.vti", out, t);
			CK(life_write_vtk(ctx, path, 1.0 /* rho_p, with Drho = 1 */, 0.0 /* ref_P */, LIFE_IO_ASYNC));
		}
	}
	snprintf(path, sizeof path, "%s/Fluid.restart", out);
	CK(life_write_restart(ctx, path, steps, LIFE_IO_ASYNC));
	CK(life_io_wait(ctx));
	printf("%lld kernel launches\n", (long long)life_launch_count(ctx));
	life_destroy(ctx);
	return 0;
}
