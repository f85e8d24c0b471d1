/* A plain C99 consumer of the drop-in boundary: nothing but include/exotica_b200_host.h, include/exotica_b200.h and the shared
 * library -- no Python, no C++ runtime in the caller, no torch types.  It does what the reference-side binding of INTEGRATION.md
 * does: load a problem from the reference's XML format (Setup::LoadProblem), read its dimensions and task indexing, and evaluate a
 * batch of configurations from host buffers (PlanningProblem::Update for every configuration).
 *
 *   consumer <package name> <package dir> <problem.xml> <n_evals>      n_evals = 0: host logic only (no CUDA call)
 *
 * Output: one line per fact, compared by tests/test_capi_host.py (host logic) and tests/test_gpu_parity.py (the batch). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "exotica_b200.h"
#include "exotica_b200_host.h"

static int fail(const char* what)
{
    fprintf(stderr, "%s: %s\n", what, exo_last_error());
    return 1;
}

/* the batch every caller of this test regenerates: a 64-bit LCG mapped into [lower, upper] of every joint */
static double next_unit(unsigned long long* s)
{
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(*s >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv)
{
    if (argc != 5)
    {
        fprintf(stderr, "usage: consumer <package> <package dir> <problem.xml> <n_evals>\n");
        return 2;
    }
    const long n_evals = atol(argv[4]);
    if (exo_host_set_package_path(argv[1], argv[2]) != 0) return fail("exo_host_set_package_path");
    FILE* f = fopen(argv[3], "rb");
    if (!f)
    {
        fprintf(stderr, "cannot open %s\n", argv[3]);
        return 2;
    }
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* xml = (char*)malloc((size_t)size + 1);
    if (fread(xml, 1, (size_t)size, f) != (size_t)size) return 2;
    xml[size] = 0;
    fclose(f);

    exo_host_problem* problem = NULL;
    if (exo_host_load_problem_xml(xml, &problem) != 0) return fail("exo_host_load_problem_xml");
    free(xml);
    int32_t dims[4];
    if (exo_host_problem_dims(problem, dims) != 0) return fail("exo_host_problem_dims");
    const int N = dims[0], LP = dims[1], LJ = dims[2];
    printf("type %s N %d length_Phi %d length_jacobian %d T %d\n", exo_host_problem_type(problem), N, LP, LJ, dims[3]);
    for (int k = 0; k < exo_host_problem_num_tasks(problem); ++k)
    {
        int32_t ix[6];
        if (exo_host_problem_task_indexing(problem, k, ix) != 0) return fail("exo_host_problem_task_indexing");
        printf("task %s id %d start %d length %d startJ %d lengthJ %d used %d\n", exo_host_problem_task_name(problem, k), ix[0], ix[1], ix[2], ix[3], ix[4], ix[5]);
    }
    exo_host_scene* scene = exo_host_problem_scene(problem);
    double* lower = (double*)malloc(sizeof(double) * (size_t)N);
    double* upper = (double*)malloc(sizeof(double) * (size_t)N);
    if (exo_host_scene_joint_limits(scene, lower, upper) != 0) return fail("exo_host_scene_joint_limits");
    printf("limits");
    for (int i = 0; i < N; ++i) printf(" %.17g %.17g", lower[i], upper[i]);
    printf("\n");
    if (n_evals > 0)
    {
        double* X = (double*)malloc(sizeof(double) * (size_t)n_evals * N);
        double* Phi = (double*)malloc(sizeof(double) * (size_t)n_evals * LP);
        double* J = (double*)malloc(sizeof(double) * (size_t)n_evals * LJ * N);
        unsigned long long s = 12345;
        for (long e = 0; e < n_evals; ++e)
            for (int i = 0; i < N; ++i) X[e * N + i] = lower[i] + (upper[i] - lower[i]) * next_unit(&s);
        /* evaluation-major host buffers in, Phi[E][length_Phi] and J[E][length_jacobian][N] out */
        if (exo_host_problem_update_batch(problem, X, n_evals, EXO_LAYOUT_AOS_ROWMAJOR, Phi, J) != 0) return fail("exo_host_problem_update_batch");
        double sp = 0.0, sj = 0.0;
        for (long k = 0; k < n_evals * LP; ++k) sp += Phi[k];
        for (long k = 0; k < n_evals * LJ * N; ++k) sj += J[k];
        printf("kernel %s launches %lld updates %lld\n", exo_problem_last_kernel(exo_host_problem_backend(problem)), (long long)exo_host_problem_launch_count(problem),
               (long long)exo_host_problem_update_count(problem));
        printf("sum_Phi %.17g sum_J %.17g\n", sp, sj);
        printf("last");
        for (int k = 0; k < LP; ++k) printf(" %.17g", Phi[(n_evals - 1) * LP + k]);
        printf("\n");
        free(X);
        free(Phi);
        free(J);
    }
    free(lower);
    free(upper);
    if (exo_host_problem_destroy(problem) != 0) return fail("exo_host_problem_destroy");
    return 0;
}
