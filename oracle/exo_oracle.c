/*
 * exo_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see exo_oracle.h for scope and citations).
 *
 * Plain C99, IEEE double, no FMA contraction (built with -ffp-contract=off), one configuration at a
 * time, the same per-element breadth-first frame update and per-frame parent-chain walks as the
 * reference.  Nothing here is shared with the CUDA product: the product compiles the tree into a
 * folded straight-line program, the oracle walks the raw element tree.
 */
#include "exo_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.1415926535897932384626433832795
#define ORC_PI_2 1.5707963267948966192313216916398

/* ------------------------------------------------------------------------------------------- */
/* orocos_kdl frames arithmetic (restated; KDL is not vendored in the reference).               */
/* ------------------------------------------------------------------------------------------- */
typedef struct
{
    double M[9]; /* KDL::Rotation::data, row-major */
    double p[3];
} frame_t;

typedef struct
{
    double vel[3];
    double rot[3];
} twist_t;

static void frame_identity(frame_t* f)
{
    memset(f, 0, sizeof(*f));
    f->M[0] = f->M[4] = f->M[8] = 1.0;
}

/* Rotation * Rotation */
static void rot_mul(const double* a, const double* b, double* out)
{
    double r[9];
    r[0] = a[0] * b[0] + a[1] * b[3] + a[2] * b[6];
    r[1] = a[0] * b[1] + a[1] * b[4] + a[2] * b[7];
    r[2] = a[0] * b[2] + a[1] * b[5] + a[2] * b[8];
    r[3] = a[3] * b[0] + a[4] * b[3] + a[5] * b[6];
    r[4] = a[3] * b[1] + a[4] * b[4] + a[5] * b[7];
    r[5] = a[3] * b[2] + a[4] * b[5] + a[5] * b[8];
    r[6] = a[6] * b[0] + a[7] * b[3] + a[8] * b[6];
    r[7] = a[6] * b[1] + a[7] * b[4] + a[8] * b[7];
    r[8] = a[6] * b[2] + a[7] * b[5] + a[8] * b[8];
    memcpy(out, r, sizeof(r));
}

/* Rotation * Vector */
static void rot_vec(const double* m, const double* v, double* out)
{
    double r0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    double r1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    double r2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    out[0] = r0;
    out[1] = r1;
    out[2] = r2;
}

/* Rotation::Inverse(Vector) = M^T v */
static void rot_inv_vec(const double* m, const double* v, double* out)
{
    double r0 = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
    double r1 = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
    double r2 = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
    out[0] = r0;
    out[1] = r1;
    out[2] = r2;
}

static void rot_transpose(const double* m, double* out)
{
    double r[9] = {m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]};
    memcpy(out, r, sizeof(r));
}

/* Frame * Frame = (M1*M2, M1*p2 + p1) */
static void frame_mul(const frame_t* a, const frame_t* b, frame_t* out)
{
    frame_t r;
    double t[3];
    rot_mul(a->M, b->M, r.M);
    rot_vec(a->M, b->p, t);
    r.p[0] = t[0] + a->p[0];
    r.p[1] = t[1] + a->p[1];
    r.p[2] = t[2] + a->p[2];
    *out = r;
}

/* Frame::Inverse = (M^T, -(M^T p)) */
static void frame_inverse(const frame_t* a, frame_t* out)
{
    frame_t r;
    double t[3];
    rot_transpose(a->M, r.M);
    rot_inv_vec(a->M, a->p, t);
    r.p[0] = -t[0];
    r.p[1] = -t[1];
    r.p[2] = -t[2];
    *out = r;
}

/* KDL cross product: Vector * Vector */
static void vec_cross(const double* a, const double* b, double* out)
{
    double r0 = a[1] * b[2] - a[2] * b[1];
    double r1 = a[2] * b[0] - a[0] * b[2];
    double r2 = a[0] * b[1] - a[1] * b[0];
    out[0] = r0;
    out[1] = r1;
    out[2] = r2;
}

/* Twist::RefPoint(d) = (vel + rot x d, rot) */
static void twist_refpoint(const twist_t* t, const double* d, twist_t* out)
{
    double c[3];
    twist_t r = *t;
    vec_cross(t->rot, d, c);
    r.vel[0] = t->vel[0] + c[0];
    r.vel[1] = t->vel[1] + c[1];
    r.vel[2] = t->vel[2] + c[2];
    *out = r;
}

/* Rotation * Twist */
static void rot_twist(const double* m, const twist_t* t, twist_t* out)
{
    twist_t r;
    rot_vec(m, t->vel, r.vel);
    rot_vec(m, t->rot, r.rot);
    *out = r;
}

/* Rotation::Rot2(unit axis, angle) */
static void rot_rot2(const double* v, double angle, double* out)
{
    double ct = cos(angle);
    double st = sin(angle);
    double vt = 1 - ct;
    double m_vt_0 = vt * v[0];
    double m_vt_1 = vt * v[1];
    double m_vt_2 = vt * v[2];
    double m_st_0 = v[0] * st;
    double m_st_1 = v[1] * st;
    double m_st_2 = v[2] * st;
    double m_vt_0_1 = m_vt_0 * v[1];
    double m_vt_0_2 = m_vt_0 * v[2];
    double m_vt_1_2 = m_vt_1 * v[2];
    out[0] = ct + m_vt_0 * v[0];
    out[1] = -m_st_2 + m_vt_0_1;
    out[2] = m_st_1 + m_vt_0_2;
    out[3] = m_st_2 + m_vt_0_1;
    out[4] = ct + m_vt_1 * v[1];
    out[5] = -m_st_0 + m_vt_1_2;
    out[6] = -m_st_1 + m_vt_0_2;
    out[7] = m_st_0 + m_vt_1_2;
    out[8] = ct + m_vt_2 * v[2];
}

static void rot_elementary(int axis /*0 x,1 y,2 z*/, double angle, double* out)
{
    double cs = cos(angle);
    double sn = sin(angle);
    memset(out, 0, 9 * sizeof(double));
    if (axis == 0)
    {
        out[0] = 1;
        out[4] = cs;
        out[5] = -sn;
        out[7] = sn;
        out[8] = cs;
    }
    else if (axis == 1)
    {
        out[0] = cs;
        out[2] = sn;
        out[4] = 1;
        out[6] = -sn;
        out[8] = cs;
    }
    else
    {
        out[0] = cs;
        out[1] = -sn;
        out[3] = sn;
        out[4] = cs;
        out[8] = 1;
    }
}

/* Rotation::Quaternion(x,y,z,w) */
static void rot_from_quaternion(double x, double y, double z, double w, double* out)
{
    double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    out[0] = w2 + x2 - y2 - z2;
    out[1] = 2 * x * y - 2 * w * z;
    out[2] = 2 * x * z + 2 * w * y;
    out[3] = 2 * x * y + 2 * w * z;
    out[4] = w2 - x2 + y2 - z2;
    out[5] = 2 * y * z - 2 * w * x;
    out[6] = 2 * x * z - 2 * w * y;
    out[7] = 2 * y * z + 2 * w * x;
    out[8] = w2 - x2 - y2 + z2;
}

/* Rotation::RPY(roll,pitch,yaw) = Rz(yaw) Ry(pitch) Rx(roll) */
static void rot_from_rpy(double roll, double pitch, double yaw, double* out)
{
    double ca1 = cos(yaw), sa1 = sin(yaw);
    double cb1 = cos(pitch), sb1 = sin(pitch);
    double cc1 = cos(roll), sc1 = sin(roll);
    out[0] = ca1 * cb1;
    out[1] = ca1 * sb1 * sc1 - sa1 * cc1;
    out[2] = ca1 * sb1 * cc1 + sa1 * sc1;
    out[3] = sa1 * cb1;
    out[4] = sa1 * sb1 * sc1 + ca1 * cc1;
    out[5] = sa1 * sb1 * cc1 - ca1 * sc1;
    out[6] = -sb1;
    out[7] = cb1 * sc1;
    out[8] = cb1 * cc1;
}

/* Vector::Norm (scaled form) */
static double vec_norm_kdl(const double* d)
{
    double tmp1 = fabs(d[0]);
    double tmp2 = fabs(d[1]);
    if (tmp1 >= tmp2)
    {
        tmp2 = fabs(d[2]);
        if (tmp1 >= tmp2)
        {
            if (tmp1 == 0) return 0;
            return tmp1 * sqrt(1 + (d[1] / d[0]) * (d[1] / d[0]) + (d[2] / d[0]) * (d[2] / d[0]));
        }
        return tmp2 * sqrt(1 + (d[0] / d[2]) * (d[0] / d[2]) + (d[1] / d[2]) * (d[1] / d[2]));
    }
    tmp1 = fabs(d[2]);
    if (tmp2 > tmp1) return tmp2 * sqrt(1 + (d[0] / d[1]) * (d[0] / d[1]) + (d[2] / d[1]) * (d[2] / d[1]));
    return tmp1 * sqrt(1 + (d[0] / d[2]) * (d[0] / d[2]) + (d[1] / d[2]) * (d[1] / d[2]));
}

static void rot_get_rpy(const double* d, double* roll, double* pitch, double* yaw)
{
    const double epsilon = 1E-12;
    *pitch = atan2(-d[6], sqrt(d[0] * d[0] + d[3] * d[3]));
    if (fabs(*pitch) > (ORC_PI_2 - epsilon))
    {
        *yaw = atan2(-d[1], d[4]);
        *roll = 0.0;
    }
    else
    {
        *roll = atan2(d[7], d[8]);
        *yaw = atan2(d[3], d[0]);
    }
}

static void rot_get_euler_zyz(const double* d, double* alpha, double* beta, double* gamma)
{
    const double epsilon = 1E-12;
    if (fabs(d[8]) > 1 - epsilon)
    {
        *gamma = 0.0;
        if (d[8] > 0)
        {
            *beta = 0.0;
            *alpha = atan2(d[3], d[0]);
        }
        else
        {
            *beta = ORC_PI;
            *alpha = atan2(-d[3], -d[0]);
        }
    }
    else
    {
        *alpha = atan2(d[5], d[2]);
        *beta = atan2(sqrt(d[6] * d[6] + d[7] * d[7]), d[8]);
        *gamma = atan2(d[7], -d[6]);
    }
}

/* Rotation::GetRotAngle (KDL 1.4 form: symmetric-matrix singularity handling, acos otherwise);
 * GetRot() = axis * angle. */
static void rot_get_rot(const double* d, double* out)
{
    const double epsilon = 1e-6;
    const double epsilon2 = 1e-6 * 10;
    double x, y, z, angle;
    if ((fabs(d[1] - d[3]) < epsilon) && (fabs(d[2] - d[6]) < epsilon) && (fabs(d[5] - d[7]) < epsilon))
    {
        if ((fabs(d[1] + d[3]) < epsilon2) && (fabs(d[2] + d[6]) < epsilon2) && (fabs(d[5] + d[7]) < epsilon2) &&
            (fabs(d[0] + d[4] + d[8] - 3) < epsilon2))
        {
            out[0] = 0.0 * 0.0;
            out[1] = 0.0 * 0.0;
            out[2] = 1.0 * 0.0;
            return;
        }
        angle = ORC_PI;
        {
            double xx = (d[0] + 1) / 2;
            double yy = (d[4] + 1) / 2;
            double zz = (d[8] + 1) / 2;
            double xy = (d[1] + d[3]) / 4;
            double xz = (d[2] + d[6]) / 4;
            double yz = (d[5] + d[7]) / 4;
            if ((xx > yy) && (xx > zz))
            {
                x = sqrt(xx);
                y = xy / x;
                z = xz / x;
            }
            else if (yy > zz)
            {
                y = sqrt(yy);
                x = xy / y;
                z = yz / y;
            }
            else
            {
                z = sqrt(zz);
                x = xz / z;
                y = yz / z;
            }
        }
        out[0] = x * angle;
        out[1] = y * angle;
        out[2] = z * angle;
        return;
    }
    {
        double f = (d[0] + d[4] + d[8] - 1) / 2;
        double n;
        if (f > 1.0) f = 1.0;
        if (f < -1.0) f = -1.0;
        angle = acos(f);
        x = (d[7] - d[5]);
        y = (d[2] - d[6]);
        z = (d[3] - d[1]);
        n = vec_norm_kdl((double[3]){x, y, z});
        /* Vector::Normalize: divides by the norm (norm cannot be ~0 here: singular cases handled above) */
        x /= n;
        y /= n;
        z /= n;
    }
    out[0] = x * angle;
    out[1] = y * angle;
    out[2] = z * angle;
}

/* Eigen::Quaterniond(Matrix3d).coeffs() -> (x,y,z,w); m(r,c) = d[3r+c] (true rotation matrix) */
static void rot_get_quaternion_eigen(const double* d, double* q)
{
    double t = d[0] + d[4] + d[8];
    if (t > 0.0)
    {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (d[7] - d[5]) * t;
        q[1] = (d[2] - d[6]) * t;
        q[2] = (d[3] - d[1]) * t;
    }
    else
    {
        int i = 0, j, k;
        if (d[4] > d[0]) i = 1;
        if (d[8] > d[4 * i]) i = 2;
        j = (i + 1) % 3;
        k = (j + 1) % 3;
        t = sqrt(d[4 * i] - d[4 * j] - d[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (d[3 * k + j] - d[3 * j + k]) * t;
        q[j] = (d[3 * j + i] + d[3 * i + j]) * t;
        q[k] = (d[3 * k + i] + d[3 * i + k]) * t;
    }
}

int exo_oracle_set_rotation(const double* d, int type, double* out)
{
    switch (type)
    {
        case EXO_ORC_QUATERNION:
            rot_get_quaternion_eigen(d, out);
            return 4;
        case EXO_ORC_RPY:
            rot_get_rpy(d, &out[0], &out[1], &out[2]);
            return 3;
        case EXO_ORC_ZYX:
            /* GetEulerZYX(a,b,c) == GetRPY(c,b,a) */
            rot_get_rpy(d, &out[2], &out[1], &out[0]);
            return 3;
        case EXO_ORC_ZYZ:
            rot_get_euler_zyz(d, &out[0], &out[1], &out[2]);
            return 3;
        case EXO_ORC_ANGLE_AXIS:
            rot_get_rot(d, out);
            return 3;
        case EXO_ORC_MATRIX:
            memcpy(out, d, 9 * sizeof(double));
            return 9;
        default:
            return -1;
    }
}

static int rotation_length(int type)
{
    static const int types[] = {4, 3, 3, 3, 3, 9};
    return (type >= 0 && type < 6) ? types[type] : -1;
}

int exo_oracle_get_frame(const double* v, int n, double* out12)
{
    frame_t f;
    frame_identity(&f);
    if (n == 7)
    {
        double norm = sqrt(v[3] * v[3] + v[4] * v[4] + v[5] * v[5] + v[6] * v[6]);
        if (norm <= 0.0) return -1;
        rot_from_quaternion(v[3] / norm, v[4] / norm, v[5] / norm, v[6] / norm, f.M);
    }
    else if (n == 6)
    {
        rot_from_rpy(v[3], v[4], v[5], f.M);
    }
    else if (n != 3)
    {
        return -1;
    }
    f.p[0] = v[0];
    f.p[1] = v[1];
    f.p[2] = v[2];
    memcpy(out12, f.M, 9 * sizeof(double));
    memcpy(out12 + 9, f.p, 3 * sizeof(double));
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Tree                                                                                          */
/* ------------------------------------------------------------------------------------------- */
typedef struct
{
    int id;
    int parent; /* element index, -1 for the root */
    int joint_type;
    double origin[3];
    double axis[3];
    frame_t f_tip;
    int control_id;
    int is_controlled;
    int is_mimic;
    int mimic_elem;
    double mimic_mult, mimic_off;
    int generated; /* slot in generated table or -1 */
    int n_children;
    int* children;
    frame_t frame;
} element_t;

typedef struct
{
    int elem_a, elem_b;
    frame_t off_a, off_b;
    frame_t temp_a, temp_b, temp_ab;
} kframe_t;

typedef struct
{
    int kind, frame_start, n_frames, rot_type, is_used;
    double param;
    int start, length, start_jacobian, length_jacobian;
    int* group;
    double* radius;
    int n_groups;
} taskmap_t;

struct exo_oracle
{
    int n_elem, n_ctrl;
    element_t* el;
    int* controlled; /* control id -> element index */
    double* tree_state; /* by element index */
    double* lower;
    double* upper;
    int n_gen;
    frame_t* gen;
    int* queue;
    /* request */
    int n_frames;
    kframe_t* fr;
    double* jac; /* [F][6*N] column-major */
    double* x;
    /* problem */
    int n_maps;
    taskmap_t* maps;
    int length_phi, length_jac;
};

exo_oracle* exo_oracle_create(int n_elem, const int* id, const int* parent, const int* joint_type,
                              const double* joint_origin, const double* joint_axis, const double* f_tip,
                              const int* control_id, const int* mimic_elem, const double* mimic_mult,
                              const double* mimic_off, const int* generated_index, int n_controlled)
{
    exo_oracle* o = (exo_oracle*)calloc(1, sizeof(exo_oracle));
    int i;
    o->n_elem = n_elem;
    o->n_ctrl = n_controlled;
    o->el = (element_t*)calloc((size_t)n_elem, sizeof(element_t));
    o->controlled = (int*)calloc((size_t)(n_controlled > 0 ? n_controlled : 1), sizeof(int));
    o->tree_state = (double*)calloc((size_t)n_elem, sizeof(double));
    o->lower = (double*)calloc((size_t)(n_controlled > 0 ? n_controlled : 1), sizeof(double));
    o->upper = (double*)calloc((size_t)(n_controlled > 0 ? n_controlled : 1), sizeof(double));
    o->queue = (int*)calloc((size_t)n_elem, sizeof(int));
    o->x = (double*)calloc((size_t)(n_controlled > 0 ? n_controlled : 1), sizeof(double));
    for (i = 0; i < n_controlled; ++i) o->controlled[i] = -1;
    for (i = 0; i < n_elem; ++i)
    {
        element_t* e = &o->el[i];
        e->id = id[i];
        e->parent = parent[i];
        e->joint_type = joint_type[i];
        memcpy(e->origin, joint_origin + 3 * i, 3 * sizeof(double));
        memcpy(e->axis, joint_axis + 3 * i, 3 * sizeof(double));
        memcpy(e->f_tip.M, f_tip + 12 * i, 9 * sizeof(double));
        memcpy(e->f_tip.p, f_tip + 12 * i + 9, 3 * sizeof(double));
        e->control_id = control_id[i];
        e->is_controlled = control_id[i] >= 0;
        e->mimic_elem = mimic_elem ? mimic_elem[i] : -1;
        e->is_mimic = e->mimic_elem >= 0;
        e->mimic_mult = mimic_mult ? mimic_mult[i] : 1.0;
        e->mimic_off = mimic_off ? mimic_off[i] : 0.0;
        e->generated = generated_index ? generated_index[i] : -1;
        if (e->generated + 1 > o->n_gen) o->n_gen = e->generated + 1;
        frame_identity(&e->frame);
        if (e->is_controlled && e->control_id < n_controlled) o->controlled[e->control_id] = i;
    }
    for (i = 0; i < n_elem; ++i)
        if (o->el[i].parent >= 0) o->el[o->el[i].parent].n_children++;
    for (i = 0; i < n_elem; ++i)
    {
        o->el[i].children = (int*)calloc((size_t)(o->el[i].n_children > 0 ? o->el[i].n_children : 1), sizeof(int));
        o->el[i].n_children = 0;
    }
    for (i = 0; i < n_elem; ++i)
        if (o->el[i].parent >= 0)
        {
            element_t* p = &o->el[o->el[i].parent];
            p->children[p->n_children++] = i;
        }
    o->gen = (frame_t*)calloc((size_t)(o->n_gen > 0 ? o->n_gen : 1), sizeof(frame_t));
    for (i = 0; i < o->n_gen; ++i) frame_identity(&o->gen[i]);
    return o;
}

static void free_maps(exo_oracle* o)
{
    int i;
    for (i = 0; i < o->n_maps; ++i)
    {
        free(o->maps[i].group);
        free(o->maps[i].radius);
    }
    free(o->maps);
    o->maps = NULL;
    o->n_maps = 0;
    o->length_phi = o->length_jac = 0;
}

void exo_oracle_destroy(exo_oracle* o)
{
    int i;
    if (!o) return;
    for (i = 0; i < o->n_elem; ++i) free(o->el[i].children);
    free(o->el);
    free(o->controlled);
    free(o->tree_state);
    free(o->lower);
    free(o->upper);
    free(o->gen);
    free(o->queue);
    free(o->fr);
    free(o->jac);
    free(o->x);
    free_maps(o);
    free(o);
}

exo_oracle* exo_oracle_clone(const exo_oracle* s)
{
    exo_oracle* o = (exo_oracle*)calloc(1, sizeof(exo_oracle));
    int i;
    int nc = s->n_ctrl > 0 ? s->n_ctrl : 1;
    *o = *s;
    o->el = (element_t*)malloc((size_t)s->n_elem * sizeof(element_t));
    memcpy(o->el, s->el, (size_t)s->n_elem * sizeof(element_t));
    for (i = 0; i < s->n_elem; ++i)
    {
        int n = s->el[i].n_children > 0 ? s->el[i].n_children : 1;
        o->el[i].children = (int*)malloc((size_t)n * sizeof(int));
        memcpy(o->el[i].children, s->el[i].children, (size_t)n * sizeof(int));
    }
#define DUP(field, count, type)                                   \
    o->field = (type*)malloc((size_t)(count) * sizeof(type));     \
    memcpy(o->field, s->field, (size_t)(count) * sizeof(type));
    DUP(controlled, nc, int)
    DUP(tree_state, s->n_elem, double)
    DUP(lower, nc, double)
    DUP(upper, nc, double)
    DUP(gen, s->n_gen > 0 ? s->n_gen : 1, frame_t)
    DUP(queue, s->n_elem, int)
    DUP(x, nc, double)
    if (s->fr)
    {
        DUP(fr, s->n_frames > 0 ? s->n_frames : 1, kframe_t)
        DUP(jac, (size_t)(s->n_frames > 0 ? s->n_frames : 1) * 6 * nc, double)
    }
    if (s->maps)
    {
        DUP(maps, s->n_maps, taskmap_t)
        for (i = 0; i < s->n_maps; ++i)
        {
            if (s->maps[i].group)
            {
                DUP(maps[i].group, s->maps[i].n_frames, int)
                DUP(maps[i].radius, s->maps[i].n_frames, double)
            }
        }
    }
#undef DUP
    return o;
}

void exo_oracle_set_tree_state(exo_oracle* o, const double* state)
{
    memcpy(o->tree_state, state, (size_t)o->n_elem * sizeof(double));
}

void exo_oracle_set_limits(exo_oracle* o, const double* lower, const double* upper)
{
    memcpy(o->lower, lower, (size_t)o->n_ctrl * sizeof(double));
    memcpy(o->upper, upper, (size_t)o->n_ctrl * sizeof(double));
}

void exo_oracle_set_generated(exo_oracle* o, int slot, const double* f)
{
    if (slot < 0 || slot >= o->n_gen) return;
    memcpy(o->gen[slot].M, f, 9 * sizeof(double));
    memcpy(o->gen[slot].p, f + 9, 3 * sizeof(double));
}

/* KDL::Joint::pose(q) */
static void joint_pose(const element_t* e, double q, frame_t* out)
{
    frame_identity(out);
    switch (e->joint_type)
    {
        case EXO_ORC_ROT_AXIS:
            rot_rot2(e->axis, q, out->M);
            memcpy(out->p, e->origin, sizeof(out->p));
            break;
        case EXO_ORC_ROT_X:
            rot_elementary(0, q, out->M);
            break;
        case EXO_ORC_ROT_Y:
            rot_elementary(1, q, out->M);
            break;
        case EXO_ORC_ROT_Z:
            rot_elementary(2, q, out->M);
            break;
        case EXO_ORC_TRANS_AXIS:
            out->p[0] = e->origin[0] + e->axis[0] * q;
            out->p[1] = e->origin[1] + e->axis[1] * q;
            out->p[2] = e->origin[2] + e->axis[2] * q;
            break;
        case EXO_ORC_TRANS_X:
            out->p[0] = q;
            break;
        case EXO_ORC_TRANS_Y:
            out->p[1] = q;
            break;
        case EXO_ORC_TRANS_Z:
            out->p[2] = q;
            break;
        default:
            break;
    }
}

/* KDL::Joint::twist(qdot = 1) */
static void joint_twist(const element_t* e, twist_t* t)
{
    memset(t, 0, sizeof(*t));
    switch (e->joint_type)
    {
        case EXO_ORC_ROT_AXIS:
            t->rot[0] = e->axis[0] * 1.0;
            t->rot[1] = e->axis[1] * 1.0;
            t->rot[2] = e->axis[2] * 1.0;
            break;
        case EXO_ORC_ROT_X:
            t->rot[0] = 1.0;
            break;
        case EXO_ORC_ROT_Y:
            t->rot[1] = 1.0;
            break;
        case EXO_ORC_ROT_Z:
            t->rot[2] = 1.0;
            break;
        case EXO_ORC_TRANS_AXIS:
            t->vel[0] = e->axis[0] * 1.0;
            t->vel[1] = e->axis[1] * 1.0;
            t->vel[2] = e->axis[2] * 1.0;
            break;
        case EXO_ORC_TRANS_X:
            t->vel[0] = 1.0;
            break;
        case EXO_ORC_TRANS_Y:
            t->vel[1] = 1.0;
            break;
        case EXO_ORC_TRANS_Z:
            t->vel[2] = 1.0;
            break;
        default:
            break;
    }
}

/* KDL::Segment::pose(q) = joint.pose(q) * f_tip */
static void segment_pose(const element_t* e, double q, frame_t* out)
{
    frame_t jp;
    joint_pose(e, q, &jp);
    frame_mul(&jp, &e->f_tip, out);
}

/* KDL::Segment::twist(q, 1) = joint.twist(1).RefPoint(joint.pose(q).M * f_tip.p) */
static void segment_twist(const element_t* e, double q, twist_t* out)
{
    frame_t jp;
    twist_t jt;
    double d[3];
    joint_pose(e, q, &jp);
    joint_twist(e, &jt);
    rot_vec(jp.M, e->f_tip.p, d);
    twist_refpoint(&jt, d, out);
}

/* KinematicElement::GetPose (kinematic_element.h:93-107) */
static void element_get_pose(const exo_oracle* o, const element_t* e, double x, frame_t* out)
{
    if (e->generated >= 0)
        *out = o->gen[e->generated];
    else if (e->is_mimic)
        segment_pose(e, e->mimic_off + e->mimic_mult * x, out);
    else
        segment_pose(e, x, out);
}

/* KinematicTree::UpdateTree (kinematic_tree.cpp:695-737): breadth-first over ALL elements */
static void update_tree(exo_oracle* o)
{
    int head = 0, tail = 0, c;
    o->queue[tail++] = 0;
    while (head < tail)
    {
        int i = o->queue[head++];
        element_t* e = &o->el[i];
        frame_t pose;
        if (e->id > -1)
        {
            if (e->joint_type != EXO_ORC_NONE)
            {
                if (!e->is_mimic)
                    element_get_pose(o, e, o->tree_state[i], &pose);
                else
                    element_get_pose(o, e, o->tree_state[e->mimic_elem], &pose);
            }
            else
            {
                element_get_pose(o, e, 0.0, &pose);
            }
            frame_mul(&o->el[e->parent].frame, &pose, &e->frame);
        }
        else
        {
            element_get_pose(o, e, 0.0, &e->frame);
        }
        for (c = 0; c < e->n_children; ++c) o->queue[tail++] = e->children[c];
    }
}

int exo_oracle_request_frames(exo_oracle* o, int n_frames, const int* elem_a, const double* off_a, const int* elem_b,
                              const double* off_b)
{
    int i;
    free(o->fr);
    free(o->jac);
    o->n_frames = n_frames;
    o->fr = (kframe_t*)calloc((size_t)(n_frames > 0 ? n_frames : 1), sizeof(kframe_t));
    o->jac = (double*)calloc((size_t)(n_frames > 0 ? n_frames : 1) * 6 * (size_t)(o->n_ctrl > 0 ? o->n_ctrl : 1), sizeof(double));
    for (i = 0; i < n_frames; ++i)
    {
        kframe_t* f = &o->fr[i];
        if (elem_a[i] < 0 || elem_a[i] >= o->n_elem || elem_b[i] < 0 || elem_b[i] >= o->n_elem) return -1;
        f->elem_a = elem_a[i];
        f->elem_b = elem_b[i];
        memcpy(f->off_a.M, off_a + 12 * i, 9 * sizeof(double));
        memcpy(f->off_a.p, off_a + 12 * i + 9, 3 * sizeof(double));
        memcpy(f->off_b.M, off_b + 12 * i, 9 * sizeof(double));
        memcpy(f->off_b.p, off_b + 12 * i + 9, 3 * sizeof(double));
    }
    return 0;
}

/* KinematicTree::FK(KinematicFrame&) (kinematic_tree.cpp:829-835) */
static void fk_frame(exo_oracle* o, kframe_t* f)
{
    frame_t inv;
    frame_mul(&o->el[f->elem_a].frame, &f->off_a, &f->temp_a);
    frame_mul(&o->el[f->elem_b].frame, &f->off_b, &f->temp_b);
    frame_inverse(&f->temp_b, &inv);
    frame_mul(&inv, &f->temp_a, &f->temp_ab);
}

/* one Jacobian column contribution (kinematic_tree.cpp:926-931) */
static void column_twist(const exo_oracle* o, const kframe_t* f, int it, twist_t* out)
{
    const element_t* e = &o->el[it];
    frame_t segment_reference;
    twist_t tw, tw2, tw3;
    double d[3], binv[9];
    if (e->parent >= 0)
        segment_reference = o->el[e->parent].frame;
    else
        frame_identity(&segment_reference);
    segment_twist(e, o->tree_state[it], &tw);
    rot_twist(segment_reference.M, &tw, &tw2);
    d[0] = f->temp_a.p[0] - e->frame.p[0];
    d[1] = f->temp_a.p[1] - e->frame.p[1];
    d[2] = f->temp_a.p[2] - e->frame.p[2];
    twist_refpoint(&tw2, d, &tw3);
    rot_transpose(f->temp_b.M, binv);
    rot_twist(binv, &tw3, out);
}

/* KinematicTree::ComputeJ (kinematic_tree.cpp:919-945) */
static void compute_j(exo_oracle* o, kframe_t* f, double* jac /* 6 x N col-major */)
{
    int it;
    memset(jac, 0, (size_t)6 * (size_t)o->n_ctrl * sizeof(double));
    fk_frame(o, f);
    it = f->elem_a;
    while (it >= 0)
    {
        const element_t* e = &o->el[it];
        if (e->is_controlled)
        {
            twist_t t;
            double* col = jac + 6 * e->control_id;
            column_twist(o, f, it, &t);
            col[0] = t.vel[0];
            col[1] = t.vel[1];
            col[2] = t.vel[2];
            col[3] = t.rot[0];
            col[4] = t.rot[1];
            col[5] = t.rot[2];
        }
        it = e->parent;
    }
    it = f->elem_b;
    while (it >= 0)
    {
        const element_t* e = &o->el[it];
        if (e->is_controlled)
        {
            twist_t t;
            double* col = jac + 6 * e->control_id;
            column_twist(o, f, it, &t);
            col[0] = col[0] - t.vel[0];
            col[1] = col[1] - t.vel[1];
            col[2] = col[2] - t.vel[2];
            col[3] = col[3] - t.rot[0];
            col[4] = col[4] - t.rot[1];
            col[5] = col[5] - t.rot[2];
        }
        it = e->parent;
    }
}

int exo_oracle_update(exo_oracle* o, const double* x)
{
    int i;
    for (i = 0; i < o->n_ctrl; ++i)
    {
        if (o->controlled[i] >= 0) o->tree_state[o->controlled[i]] = x[i];
        o->x[i] = x[i];
    }
    update_tree(o);
    for (i = 0; i < o->n_frames; ++i) fk_frame(o, &o->fr[i]);                                       /* UpdateFK */
    for (i = 0; i < o->n_frames; ++i) compute_j(o, &o->fr[i], o->jac + (size_t)i * 6 * o->n_ctrl); /* UpdateJ */
    return 0;
}

void exo_oracle_get_phi(const exo_oracle* o, double* out)
{
    int i;
    for (i = 0; i < o->n_frames; ++i)
    {
        memcpy(out + 12 * i, o->fr[i].temp_ab.M, 9 * sizeof(double));
        memcpy(out + 12 * i + 9, o->fr[i].temp_ab.p, 3 * sizeof(double));
    }
}

void exo_oracle_get_jacobian(const exo_oracle* o, double* out)
{
    memcpy(out, o->jac, (size_t)o->n_frames * 6 * (size_t)o->n_ctrl * sizeof(double));
}

void exo_oracle_get_element_frame(const exo_oracle* o, int elem, double* out12)
{
    memcpy(out12, o->el[elem].frame.M, 9 * sizeof(double));
    memcpy(out12 + 9, o->el[elem].frame.p, 3 * sizeof(double));
}

/* ------------------------------------------------------------------------------------------- */
/* Task maps and the problem-level Update                                                        */
/* ------------------------------------------------------------------------------------------- */
void exo_oracle_problem_clear(exo_oracle* o)
{
    free_maps(o);
}

int exo_oracle_problem_add_map(exo_oracle* o, int kind, int frame_start, int n_frames, int rot_type, double param,
                               const int* sphere_group, const double* sphere_radius, int is_used)
{
    taskmap_t* m;
    int i;
    o->maps = (taskmap_t*)realloc(o->maps, (size_t)(o->n_maps + 1) * sizeof(taskmap_t));
    m = &o->maps[o->n_maps];
    memset(m, 0, sizeof(*m));
    m->kind = kind;
    m->frame_start = frame_start;
    m->n_frames = n_frames;
    m->rot_type = rot_type;
    m->param = param;
    m->is_used = is_used;
    switch (kind)
    {
        case EXO_ORC_MAP_EFF_FRAME:
            if (rotation_length(rot_type) < 0) return -1;
            m->length = n_frames * (3 + rotation_length(rot_type)); /* eff_frame.cpp:92-95 */
            m->length_jacobian = n_frames * 6;                      /* eff_frame.cpp:97-100 */
            break;
        case EXO_ORC_MAP_EFF_POSITION:
            m->length = m->length_jacobian = n_frames * 3; /* eff_position.cpp:72-75 */
            break;
        case EXO_ORC_MAP_JOINT_LIMIT:
            m->length = m->length_jacobian = o->n_ctrl; /* joint_limit.cpp:58-61 */
            break;
        case EXO_ORC_MAP_EFF_ORIENTATION:
            if (rotation_length(rot_type) < 0) return -1;
            m->length = n_frames * rotation_length(rot_type); /* eff_orientation.cpp:100-103 */
            m->length_jacobian = n_frames * 3;                /* eff_orientation.cpp:105-108 */
            break;
        case EXO_ORC_MAP_EFF_POSITION_XY:
            m->length = m->length_jacobian = n_frames * 2;
            break;
        case EXO_ORC_MAP_DISTANCE:
            m->length = m->length_jacobian = n_frames;
            break;
        case EXO_ORC_MAP_JOINT_POSE:
        {
            /* joint_pose.cpp:66-97: sphere_group carries JointMap, sphere_radius JointRef, param their length;
             * no frames (n_frames must be 0) */
            const int n = (int)param;
            m->n_groups = n;
            m->group = (int*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
            m->radius = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
            for (i = 0; i < n; ++i)
            {
                m->group[i] = sphere_group[i];
                m->radius[i] = sphere_radius[i];
            }
            m->length = m->length_jacobian = n;
            break;
        }
        case EXO_ORC_MAP_EFF_AXIS_ALIGNMENT:
            /* 2n frames; sphere_radius carries the n directions (3 doubles each) */
            m->radius = (double*)malloc((size_t)(n_frames > 0 ? 3 * (n_frames / 2) : 1) * sizeof(double));
            for (i = 0; i < 3 * (n_frames / 2); ++i) m->radius[i] = sphere_radius[i];
            m->length = m->length_jacobian = n_frames / 2; /* eff_axis_alignment.cpp:171-174 */
            break;
        case EXO_ORC_MAP_CENTER_OF_MASS:
            /* sphere_radius carries the masses, param dim_ = EnableZ ? 3 : 2 (center_of_mass.cpp:134-137,196) */
            m->radius = (double*)malloc((size_t)(n_frames > 0 ? n_frames : 1) * sizeof(double));
            for (i = 0; i < n_frames; ++i) m->radius[i] = sphere_radius[i];
            m->length = m->length_jacobian = (int)param;
            break;
        case EXO_ORC_MAP_SPHERE_COLLISION:
            m->group = (int*)malloc((size_t)(n_frames > 0 ? n_frames : 1) * sizeof(int));
            m->radius = (double*)malloc((size_t)(n_frames > 0 ? n_frames : 1) * sizeof(double));
            m->n_groups = 0;
            for (i = 0; i < n_frames; ++i)
            {
                m->group[i] = sphere_group[i];
                m->radius[i] = sphere_radius[i];
                if (sphere_group[i] + 1 > m->n_groups) m->n_groups = sphere_group[i] + 1;
            }
            m->length = m->length_jacobian = m->n_groups * (m->n_groups - 1) / 2; /* sphere_collision.cpp:69 */
            break;
        default:
            return -1;
    }
    m->start = o->length_phi;
    m->start_jacobian = o->length_jac;
    o->length_phi += m->length;
    o->length_jac += m->length_jacobian;
    o->n_maps++;
    return o->n_maps - 1;
}

int exo_oracle_problem_length_phi(const exo_oracle* o)
{
    return o->length_phi;
}

int exo_oracle_problem_length_jacobian(const exo_oracle* o)
{
    return o->length_jac;
}

static void map_eff_frame(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    int i, r, c;
    int small = rotation_length(m->rot_type);
    int big = small + 3;
    for (i = 0; i < m->n_frames; ++i)
    {
        const kframe_t* f = &o->fr[m->frame_start + i];
        const double* kj = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
        phi[i * big + 0] = f->temp_ab.p[0];
        phi[i * big + 1] = f->temp_ab.p[1];
        phi[i * big + 2] = f->temp_ab.p[2];
        exo_oracle_set_rotation(f->temp_ab.M, m->rot_type, phi + i * big + 3);
        if (jac)
            for (c = 0; c < o->n_ctrl; ++c)
                for (r = 0; r < 6; ++r) jac[(size_t)c * ldj + i * 6 + r] = kj[6 * c + r];
    }
}

static void map_eff_position(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    int i, r, c;
    for (i = 0; i < m->n_frames; ++i)
    {
        const kframe_t* f = &o->fr[m->frame_start + i];
        const double* kj = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
        phi[i * 3 + 0] = f->temp_ab.p[0];
        phi[i * 3 + 1] = f->temp_ab.p[1];
        phi[i * 3 + 2] = f->temp_ab.p[2];
        if (jac)
            for (c = 0; c < o->n_ctrl; ++c)
                for (r = 0; r < 3; ++r) jac[(size_t)c * ldj + i * 3 + r] = kj[6 * c + r];
    }
}

/* EffOrientation::Update (eff_orientation.cpp:53-71) */
static void map_eff_orientation(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    int i, r, c;
    const int stride = rotation_length(m->rot_type);
    for (i = 0; i < m->n_frames; ++i)
    {
        const kframe_t* f = &o->fr[m->frame_start + i];
        const double* kj = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
        exo_oracle_set_rotation(f->temp_ab.M, m->rot_type, phi + i * stride);
        if (jac)
            for (c = 0; c < o->n_ctrl; ++c)
                for (r = 0; r < 3; ++r) jac[(size_t)c * ldj + i * 3 + r] = kj[6 * c + 3 + r]; /* bottomRows<3> */
    }
}

/* EffPositionXY::Update (eff_position_xy.cpp:36-57) */
static void map_eff_position_xy(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    int i, r, c;
    for (i = 0; i < m->n_frames; ++i)
    {
        const kframe_t* f = &o->fr[m->frame_start + i];
        const double* kj = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
        phi[i * 2 + 0] = f->temp_ab.p[0];
        phi[i * 2 + 1] = f->temp_ab.p[1];
        if (jac)
            for (c = 0; c < o->n_ctrl; ++c)
                for (r = 0; r < 2; ++r) jac[(size_t)c * ldj + i * 2 + r] = kj[6 * c + r];
    }
}

/* Distance::Update (distance.cpp:36-57) */
static void map_distance(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    int i, c;
    for (i = 0; i < m->n_frames; ++i)
    {
        const kframe_t* f = &o->fr[m->frame_start + i];
        const double* kj = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
        const double* p = f->temp_ab.p;
        phi[i] = vec_norm_kdl(p);
        if (jac)
            for (c = 0; c < o->n_ctrl; ++c) jac[(size_t)c * ldj + i] = (p[0] * kj[6 * c] + p[1] * kj[6 * c + 1] + p[2] * kj[6 * c + 2]) / phi[i];
    }
}

/* EffAxisAlignment::Update (eff_axis_alignment.cpp:116-145): the map holds 2n frames, frame i + n being frame i with
 * frame_A_offset.p replaced by the (normalised) axis (:53-57); m->radius carries the n normalised directions (3 each). */
static void map_eff_axis_alignment(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    const int n = m->n_frames / 2;
    int i, c, k;
    for (i = 0; i < n; ++i)
    {
        const kframe_t* f0 = &o->fr[m->frame_start + i];
        const kframe_t* f1 = &o->fr[m->frame_start + i + n];
        const double* j0 = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
        const double* j1 = o->jac + (size_t)(m->frame_start + i + n) * 6 * o->n_ctrl;
        const double* dir = m->radius + 3 * i;
        double a[3];
        for (k = 0; k < 3; ++k) a[k] = f1->temp_ab.p[k] - f0->temp_ab.p[k];
        phi[i] = (a[0] * dir[0] + a[1] * dir[1] + a[2] * dir[2]) - 1.0;
        if (jac)
            for (c = 0; c < o->n_ctrl; ++c)
                jac[(size_t)c * ldj + i] = dir[0] * (j1[6 * c] - j0[6 * c]) + dir[1] * (j1[6 * c + 1] - j0[6 * c + 1]) + dir[2] * (j1[6 * c + 2] - j0[6 * c + 2]);
    }
}

/* CenterOfMass::Update with explicit frames (center_of_mass.cpp:37-54, 56-86): com = sum_i p_i mass_i / M,
 * jacobian += mass_i / M * J_i.topRows(dim); m->radius carries the masses, m->param dim_ (3, or 2 without z).
 * M == 0 returns before anything is written. */
static void map_center_of_mass(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    const int dim = (int)m->param;
    double M = 0.0, com[3] = {0.0, 0.0, 0.0};
    int i, c, k;
    for (i = 0; i < m->n_frames; ++i) M += m->radius[i];
    if (M == 0.0) return;
    for (i = 0; i < m->n_frames; ++i)
    {
        const kframe_t* f = &o->fr[m->frame_start + i];
        const double* kj = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
        for (k = 0; k < 3; ++k) com[k] += f->temp_ab.p[k] * m->radius[i];
        if (jac)
            for (c = 0; c < o->n_ctrl; ++c)
                for (k = 0; k < dim; ++k) jac[(size_t)c * ldj + k] += m->radius[i] / M * kj[6 * c + k];
    }
    for (k = 0; k < dim; ++k) phi[k] = com[k] / M;
}

/* JointPose::Update (joint_pose.cpp:36-57) */
static void map_joint_pose(const exo_oracle* o, const taskmap_t* m, const double* q, double* phi, double* jac, int ldj)
{
    int i;
    for (i = 0; i < m->n_groups; ++i)
    {
        phi[i] = q[m->group[i]] - m->radius[i];
        if (jac) jac[(size_t)m->group[i] * ldj + i] = 1.0;
    }
}

static void map_joint_limit(const exo_oracle* o, const taskmap_t* m, const double* q, double* phi, double* jac, int ldj)
{
    int i;
    const double safe = m->param;
    for (i = 0; i < o->n_ctrl; ++i)
    {
        const double lo = o->lower[i], hi = o->upper[i];
        const double tau = (0.5 * safe) * (hi - lo);
        /* joint_limit.cpp:72-75: Eigen select; the upper test is applied second */
        if (q[i] < lo + tau) phi[i] = q[i] - lo - tau;
        if (q[i] > hi - tau) phi[i] = q[i] - hi + tau;
        if (jac)
        {
            if (q[i] >= lo + tau && q[i] <= hi - tau)
                jac[(size_t)i * ldj + i] = 0;
            else
                jac[(size_t)i * ldj + i] = 1;
        }
    }
}

static void map_sphere_collision(const exo_oracle* o, const taskmap_t* m, double* phi, double* jac, int ldj)
{
    const double eps = 1.0 / m->param; /* sphere_collision.cpp:39 */
    int ga, gb, i, j, c, phi_i = 0;
    for (i = 0; i < m->length; ++i) phi[i] = 0.0;
    if (jac)
        for (c = 0; c < o->n_ctrl; ++c)
            for (i = 0; i < m->length; ++i) jac[(size_t)c * ldj + i] = 0.0;
    for (ga = 0; ga < m->n_groups; ++ga)
    {
        for (gb = ga + 1; gb < m->n_groups; ++gb)
        {
            for (i = 0; i < m->n_frames; ++i)
            {
                if (m->group[i] != ga) continue;
                for (j = 0; j < m->n_frames; ++j)
                {
                    const kframe_t *fa, *fb;
                    double d[3];
                    if (m->group[j] != gb) continue;
                    fa = &o->fr[m->frame_start + i];
                    fb = &o->fr[m->frame_start + j];
                    d[0] = fa->temp_ab.p[0] - fb->temp_ab.p[0];
                    d[1] = fa->temp_ab.p[1] - fb->temp_ab.p[1];
                    d[2] = fa->temp_ab.p[2] - fb->temp_ab.p[2];
                    /* Distance(): sphere_collision.cpp:72-75 */
                    phi[phi_i] += 1.0 / (1.0 + exp(5.0 * eps * (vec_norm_kdl(d) - m->radius[i] - m->radius[j])));
                    if (jac)
                    {
                        /* Jacobian(): sphere_collision.cpp:77-85 */
                        const double* ja = o->jac + (size_t)(m->frame_start + i) * 6 * o->n_ctrl;
                        const double* jb = o->jac + (size_t)(m->frame_start + j) * 6 * o->n_ctrl;
                        const double norm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                        for (c = 0; c < o->n_ctrl; ++c)
                        {
                            double dot = (ja[6 * c + 0] - jb[6 * c + 0]) * d[0] + (ja[6 * c + 1] - jb[6 * c + 1]) * d[1] +
                                         (ja[6 * c + 2] - jb[6 * c + 2]) * d[2];
                            jac[(size_t)c * ldj + phi_i] += -dot / norm;
                        }
                    }
                }
            }
            ++phi_i;
        }
    }
}

int exo_oracle_problem_update(exo_oracle* o, const double* x, double* phi, double* jac)
{
    int i, k;
    const int ldj = o->length_jac;
    exo_oracle_update(o, x); /* scene_->Update(x, t) */
    /* Phi.SetZero(length_Phi): zeros, then identity rotation encodings at the Lie-group entries
     * (task_space_vector.cpp:54-62; entries from EffFrame::GetLieGroupIndices eff_frame.cpp:43-51) */
    for (i = 0; i < o->length_phi; ++i) phi[i] = 0.0;
    for (k = 0; k < o->n_maps; ++k)
    {
        const taskmap_t* m = &o->maps[k];
        if (m->kind == EXO_ORC_MAP_EFF_FRAME || m->kind == EXO_ORC_MAP_EFF_ORIENTATION)
        {
            /* GetLieGroupIndices: eff_frame.cpp:43-51, eff_orientation.cpp:42-51 */
            static const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            const int lead = m->kind == EXO_ORC_MAP_EFF_FRAME ? 3 : 0;
            int big = lead + rotation_length(m->rot_type);
            for (i = 0; i < m->n_frames; ++i) exo_oracle_set_rotation(ident, m->rot_type, phi + m->start + i * big + lead);
        }
    }
    if (jac) memset(jac, 0, (size_t)ldj * (size_t)o->n_ctrl * sizeof(double)); /* jacobian.setZero() */
    for (k = 0; k < o->n_maps; ++k)
    {
        const taskmap_t* m = &o->maps[k];
        double* jrows = jac ? jac + m->start_jacobian : NULL;
        if (!m->is_used) continue; /* abstract_time_indexed_problem.cpp:250 */
        switch (m->kind)
        {
            case EXO_ORC_MAP_EFF_FRAME:
                map_eff_frame(o, m, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_EFF_POSITION:
                map_eff_position(o, m, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_JOINT_LIMIT:
                map_joint_limit(o, m, x, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_SPHERE_COLLISION:
                map_sphere_collision(o, m, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_EFF_ORIENTATION:
                map_eff_orientation(o, m, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_EFF_POSITION_XY:
                map_eff_position_xy(o, m, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_DISTANCE:
                map_distance(o, m, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_JOINT_POSE:
                map_joint_pose(o, m, x, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_EFF_AXIS_ALIGNMENT:
                map_eff_axis_alignment(o, m, phi + m->start, jrows, ldj);
                break;
            case EXO_ORC_MAP_CENTER_OF_MASS:
                map_center_of_mass(o, m, phi + m->start, jrows, ldj);
                break;
            default:
                return -1;
        }
    }
    return 0;
}

typedef struct
{
    exo_oracle* o;
    const double* X;
    long e0, e1;
    int T;
    const double* gen;
    int n_gen;
    double* phi;
    double* jac;
} chunk_t;

static void* run_chunk(void* arg)
{
    chunk_t* c = (chunk_t*)arg;
    exo_oracle* o = c->o;
    const size_t lp = (size_t)o->length_phi, lj = (size_t)o->length_jac * (size_t)o->n_ctrl;
    double* tmp_phi = (double*)malloc((lp > 0 ? lp : 1) * sizeof(double));
    double* tmp_jac = (double*)malloc((lj > 0 ? lj : 1) * sizeof(double));
    long e;
    int g;
    for (e = c->e0; e < c->e1; ++e)
    {
        if (c->gen)
        {
            int t = (int)(e % c->T);
            for (g = 0; g < c->n_gen; ++g) exo_oracle_set_generated(o, g, c->gen + ((size_t)t * c->n_gen + g) * 12);
        }
        exo_oracle_problem_update(o, c->X + (size_t)e * o->n_ctrl, c->phi ? c->phi + (size_t)e * lp : tmp_phi,
                                  c->jac ? c->jac + (size_t)e * lj : tmp_jac);
    }
    free(tmp_phi);
    free(tmp_jac);
    return NULL;
}

int exo_oracle_problem_update_batch(exo_oracle* o, const double* X, long E, int T, const double* gen, int n_gen,
                                    double* phi, double* jac, int n_threads)
{
    int i;
    if (T < 1) T = 1;
    if (n_threads <= 1)
    {
        chunk_t c = {o, X, 0, E, T, gen, n_gen, phi, jac};
        run_chunk(&c);
        return 0;
    }
    {
        pthread_t* th = (pthread_t*)malloc((size_t)n_threads * sizeof(pthread_t));
        chunk_t* ch = (chunk_t*)malloc((size_t)n_threads * sizeof(chunk_t));
        long per = (E + n_threads - 1) / n_threads;
        for (i = 0; i < n_threads; ++i)
        {
            long e0 = per * i, e1 = per * (i + 1);
            if (e0 > E) e0 = E;
            if (e1 > E) e1 = E;
            ch[i].o = exo_oracle_clone(o); /* the tree is stateful: one instance per thread */
            ch[i].X = X;
            ch[i].e0 = e0;
            ch[i].e1 = e1;
            ch[i].T = T;
            ch[i].gen = gen;
            ch[i].n_gen = n_gen;
            ch[i].phi = phi;
            ch[i].jac = jac;
            pthread_create(&th[i], NULL, run_chunk, &ch[i]);
        }
        for (i = 0; i < n_threads; ++i)
        {
            pthread_join(th[i], NULL);
            exo_oracle_destroy(ch[i].o);
        }
        free(th);
        free(ch);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Task cost: TaskSpaceVector difference, scalar task and transition cost */
/* ------------------------------------------------------------------------------------------- */
/* KDL::Rotation::EulerZYZ(alfa, beta, gamma) */
static void rot_from_euler_zyz(double a, double b, double g, double* out)
{
    double sa = sin(a), ca = cos(a), sb = sin(b), cb = cos(b), sg = sin(g), cg = cos(g);
    out[0] = ca * cb * cg - sa * sg;
    out[1] = -ca * cb * sg - sa * cg;
    out[2] = ca * sb;
    out[3] = sa * cb * cg + ca * sg;
    out[4] = -sa * cb * sg + ca * cg;
    out[5] = sa * sb;
    out[6] = -sb * cg;
    out[7] = sb * sg;
    out[8] = cb;
}

/* exotica::GetRotation (exotica_core/src/tools/conversions.cpp:114-148) */
int exo_oracle_get_rotation(const double* data, int type, double* out)
{
    static const double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    switch (type)
    {
        case EXO_ORC_QUATERNION:
            rot_from_quaternion(data[0], data[1], data[2], data[3], out); /* KDL::Rotation::Quaternion: no normalisation */
            return 0;
        case EXO_ORC_RPY:
            rot_from_rpy(data[0], data[1], data[2], out);
            return 0;
        case EXO_ORC_ZYX:
            rot_from_rpy(data[2], data[1], data[0], out); /* EulerZYX(a,b,c) = RPY(c,b,a) */
            return 0;
        case EXO_ORC_ZYZ:
            rot_from_euler_zyz(data[0], data[1], data[2], out);
            return 0;
        case EXO_ORC_ANGLE_AXIS:
        {
            double angle = vec_norm_kdl(data);
            if (fabs(angle) > 1e-10)
            {
                /* KDL::Rotation::Rot(axis, angle): normalises the axis, then Rot2 */
                double ax[3] = {data[0] / angle, data[1] / angle, data[2] / angle};
                rot_rot2(ax, angle, out);
            }
            else
                memcpy(out, ident, sizeof(ident));
            return 0;
        }
        case EXO_ORC_MATRIX:
            memcpy(out, data, 9 * sizeof(double));
            return 0;
        default:
            return -1;
    }
}

/* TaskSpaceVector::operator- (exotica_core/src/task_space_vector.cpp:64-91) on the problem's Phi layout:
 * ydiff[length_jacobian]; rotation sub-blocks become rotvec = M1 * (M2^-1 M1).GetRot(). */
int exo_oracle_problem_ydiff(const exo_oracle* o, const double* phi, const double* goal, double* ydiff)
{
    int k, i, f;
    for (k = 0; k < o->n_maps; ++k)
    {
        const taskmap_t* m = &o->maps[k];
        if (m->kind == EXO_ORC_MAP_EFF_FRAME || m->kind == EXO_ORC_MAP_EFF_ORIENTATION)
        {
            const int lead = m->kind == EXO_ORC_MAP_EFF_FRAME ? 3 : 0;
            const int rl = rotation_length(m->rot_type), big = lead + rl;
            for (f = 0; f < m->n_frames; ++f)
            {
                const double* a = phi + m->start + f * big;
                const double* b = goal + m->start + f * big;
                double* out = ydiff + m->start_jacobian + (lead + 3) * f;
                double M1[9], M2[9], M2t[9], M[9], rv[3];
                for (i = 0; i < lead; ++i) out[i] = a[i] - b[i];
                if (exo_oracle_get_rotation(a + lead, m->rot_type, M1) || exo_oracle_get_rotation(b + lead, m->rot_type, M2)) return -1;
                rot_transpose(M2, M2t);
                rot_mul(M2t, M1, M);
                rot_get_rot(M, rv);
                rot_vec(M1, rv, out + lead);
            }
        }
        else
            for (i = 0; i < m->length; ++i) ydiff[m->start_jacobian + i] = phi[m->start + i] - goal[m->start + i];
    }
    return 0;
}

/* Scalar cost of one trajectory X[T][N] (AbstractTimeIndexedProblem::GetCost, abstract_time_indexed_problem.cpp:299-342):
 *   sum_{t=t_first}^{T-1} ct * ( ydiff[t]^T S[t] ydiff[t] + xdiff[t]^T W xdiff[t] ),  S[t] = diag(rho[t][map] per Jacobian row),
 * xdiff[t] = x[t] - x[t-1].  End-pose problems: T = 1, t_first = 0, ct = 1, W = NULL
 * (UnconstrainedEndPoseProblem::GetScalarCost, unconstrained_end_pose_problem.cpp:79-82).
 * rho[T][n_maps]; goal[T_goal][length_Phi] with T_goal in {1, T}; gen: per-timestep generated frames or NULL. */
double exo_oracle_trajectory_cost(exo_oracle* o, const double* X, int T, int t_first, const double* rho, const double* goal, int T_goal,
                                  double ct, const double* W, const double* gen, int n_gen, double* cost_t)
{
    const int N = o->n_ctrl, lp = o->length_phi, lj = o->length_jac;
    double* phi = (double*)malloc((size_t)(lp > 0 ? lp : 1) * sizeof(double));
    double* yd = (double*)malloc((size_t)(lj > 0 ? lj : 1) * sizeof(double));
    double total = 0.0;
    int t, k, i, g;
    for (t = 0; t < T; ++t)
    {
        double c = 0.0;
        if (cost_t) cost_t[t] = 0.0;
        if (t < t_first) continue;
        if (gen)
            for (g = 0; g < n_gen; ++g) exo_oracle_set_generated(o, g, gen + ((size_t)t * n_gen + g) * 12);
        exo_oracle_problem_update(o, X + (size_t)t * N, phi, NULL);
        exo_oracle_problem_ydiff(o, phi, goal + (size_t)(T_goal > 1 ? t : 0) * lp, yd);
        for (k = 0; k < o->n_maps; ++k)
        {
            const taskmap_t* m = &o->maps[k];
            double s = 0.0;
            for (i = 0; i < m->length_jacobian; ++i) s += yd[m->start_jacobian + i] * yd[m->start_jacobian + i];
            c += rho[(size_t)t * o->n_maps + k] * s;
        }
        if (W && t > 0)
            for (i = 0; i < N; ++i)
            {
                double d = X[(size_t)t * N + i] - X[(size_t)(t - 1) * N + i];
                c += W[i] * d * d;
            }
        c *= ct;
        if (cost_t) cost_t[t] = c;
        total += c;
    }
    free(phi);
    free(yd);
    return total;
}

/* AbstractTimeIndexedProblem::GetCostJacobian (abstract_time_indexed_problem.cpp:309-318) with GetScalarTaskJacobian =
 * 2 ct J^T S ydiff (:326-330) and GetScalarTransitionJacobian = 2 ct W xdiff (:338-342):
 *   grad[t] = task(t) + transition(t) - transition(t + 1)    for t = 1 .. T-1  (the last timestep has no successor),
 * written to grad[T][N] with row 0 left at zero (the reference's vector starts at t = 1).  End-pose problems
 * (UnconstrainedEndPoseProblem::GetScalarJacobian, unconstrained_end_pose_problem.cpp:84-87): T = 1, t_first = 0,
 * ct = 1, W = NULL -> grad[0] = 2 J^T S ydiff. */
void exo_oracle_trajectory_cost_jacobian(exo_oracle* o, const double* X, int T, int t_first, const double* rho, const double* goal, int T_goal,
                                         double ct, const double* W, const double* gen, int n_gen, double* grad)
{
    const int N = o->n_ctrl, lp = o->length_phi, lj = o->length_jac;
    double* phi = (double*)malloc((size_t)(lp > 0 ? lp : 1) * sizeof(double));
    double* jac = (double*)malloc((size_t)(lj > 0 ? lj : 1) * N * sizeof(double));
    double* yd = (double*)malloc((size_t)(lj > 0 ? lj : 1) * sizeof(double));
    int t, k, i, g, n;
    for (i = 0; i < T * N; ++i) grad[i] = 0.0;
    for (t = t_first; t < T; ++t)
    {
        double* gt = grad + (size_t)t * N;
        if (gen)
            for (g = 0; g < n_gen; ++g) exo_oracle_set_generated(o, g, gen + ((size_t)t * n_gen + g) * 12);
        exo_oracle_problem_update(o, X + (size_t)t * N, phi, jac);
        exo_oracle_problem_ydiff(o, phi, goal + (size_t)(T_goal > 1 ? t : 0) * lp, yd);
        for (k = 0; k < o->n_maps; ++k)
        {
            const taskmap_t* m = &o->maps[k];
            const double r = rho[(size_t)t * o->n_maps + k];
            for (i = 0; i < m->length_jacobian; ++i)
                for (n = 0; n < N; ++n) gt[n] += jac[(size_t)n * lj + (m->start_jacobian + i)] * (r * yd[m->start_jacobian + i]);  /* column-major J */
        }
        for (n = 0; n < N; ++n) gt[n] *= 2.0 * ct;
        if (W && t > 0)
            for (n = 0; n < N; ++n)
            {
                const double tr = 2.0 * ct * W[n] * (X[(size_t)t * N + n] - X[(size_t)(t - 1) * N + n]);
                gt[n] += tr;
                if (t > 1) grad[(size_t)(t - 1) * N + n] -= tr;
            }
    }
    free(phi);
    free(jac);
    free(yd);
}

/* AICOSolver::GetTaskCosts (exotations/solvers/exotica_aico_solver/src/aico_solver.cpp:364-386) after prob_->Update(qhat[t], t)
 * for every timestep of one trajectory: for each task with precision rho > 0
 *   C += rho |ydiff_i|^2,  R += rho J_i^T J_i,  v = -ydiff_i + J_i qhat,  r += rho J_i^T v,  rhat += rho |v|^2,
 * cost[t] = ct * C.  R[T][N][N] row-major (symmetric), r[T][N], rhat[T], cost[T]. */
void exo_oracle_aico_task_messages(exo_oracle* o, const double* Q, int T, const double* rho, const double* goal, int T_goal, double ct,
                                   const double* gen, int n_gen, double* R, double* r, double* rhat, double* cost)
{
    const int N = o->n_ctrl, lp = o->length_phi, lj = o->length_jac;
    double* phi = (double*)malloc((size_t)(lp > 0 ? lp : 1) * sizeof(double));
    double* jac = (double*)malloc((size_t)(lj > 0 ? lj : 1) * N * sizeof(double));
    double* yd = (double*)malloc((size_t)(lj > 0 ? lj : 1) * sizeof(double));
    int t, k, i, a, b, g;
    for (t = 0; t < T; ++t)
    {
        const double* q = Q + (size_t)t * N;
        double* Rt = R + (size_t)t * N * N;
        double* rt = r + (size_t)t * N;
        double C = 0.0;
        if (gen)
            for (g = 0; g < n_gen; ++g) exo_oracle_set_generated(o, g, gen + ((size_t)t * n_gen + g) * 12);
        exo_oracle_problem_update(o, q, phi, jac);
        exo_oracle_problem_ydiff(o, phi, goal + (size_t)(T_goal > 1 ? t : 0) * lp, yd);
        for (i = 0; i < N * N; ++i) Rt[i] = 0.0;
        for (i = 0; i < N; ++i) rt[i] = 0.0;
        rhat[t] = 0.0;
        for (k = 0; k < o->n_maps; ++k)
        {
            const taskmap_t* m = &o->maps[k];
            const double prec = rho[(size_t)t * o->n_maps + k];
            if (!(prec > 0)) continue;
            for (i = 0; i < m->length_jacobian; ++i)
            {
                const int row = m->start_jacobian + i;
                double v = -yd[row];
                for (a = 0; a < N; ++a) v += jac[(size_t)a * lj + row] * q[a];
                C += prec * yd[row] * yd[row];
                rhat[t] += prec * v * v;
                for (a = 0; a < N; ++a)
                {
                    rt[a] += prec * jac[(size_t)a * lj + row] * v;
                    for (b = 0; b < N; ++b) Rt[(size_t)a * N + b] += prec * jac[(size_t)a * lj + row] * jac[(size_t)b * lj + row];
                }
            }
        }
        cost[t] = ct * C;
    }
    free(phi);
    free(jac);
    free(yd);
}

/* ------------------------------------------------------------------------------------------- */
/* Task-map Hessians                                                    */
/* ------------------------------------------------------------------------------------------- */
/* KinematicTree::ComputeH (kinematic_tree.cpp:947-978): hess[6][N][N] (row-major N x N each) from a 6 x N column-major
 * Jacobian; Twist * Twist with a zero linear part on the left: (rot x vel, rot x rot). */
static void compute_h(const double* kj, int N, double* hess)
{
    int i, j, k;
    memset(hess, 0, (size_t)6 * N * N * sizeof(double));
    for (i = 0; i < N; ++i)
    {
        const double* wi = kj + 6 * i + 3;
        for (j = i; j < N; ++j)
        {
            const double* vj = kj + 6 * j;
            const double* wj = kj + 6 * j + 3;
            double Hij[6];
            vec_cross(wi, vj, Hij);
            vec_cross(wi, wj, Hij + 3);
            for (k = 0; k < 3; ++k)
            {
                hess[(size_t)k * N * N + (size_t)i * N + j] = Hij[k];
                hess[(size_t)k * N * N + (size_t)j * N + i] = Hij[k];
            }
            if (i != j)
                for (k = 3; k < 6; ++k) hess[(size_t)k * N * N + (size_t)j * N + i] = Hij[k];
        }
    }
}

/* Problem-level Hessian of one configuration: hess[length_jacobian][N][N].  Runs exo_oracle_problem_update first. */
int exo_oracle_problem_hessian(exo_oracle* o, const double* x, double* hess)
{
    const int N = o->n_ctrl, lp = o->length_phi, lj = o->length_jac;
    const size_t NN = (size_t)N * N;
    double* phi = (double*)malloc((size_t)(lp > 0 ? lp : 1) * sizeof(double));
    double* jac = (double*)malloc((size_t)(lj > 0 ? lj : 1) * N * sizeof(double));
    double* fh = (double*)malloc(6 * NN * sizeof(double));
    int k, f, r, i, j;
    exo_oracle_problem_update(o, x, phi, jac);
    memset(hess, 0, (size_t)lj * NN * sizeof(double));
    for (k = 0; k < o->n_maps; ++k)
    {
        const taskmap_t* m = &o->maps[k];
        double* hm = hess + (size_t)m->start_jacobian * NN;
        if (!m->is_used || m->kind == EXO_ORC_MAP_JOINT_POSE) continue; /* zero */
        if (m->kind == EXO_ORC_MAP_JOINT_LIMIT || m->kind == EXO_ORC_MAP_SPHERE_COLLISION || m->kind == EXO_ORC_MAP_EFF_AXIS_ALIGNMENT ||
            m->kind == EXO_ORC_MAP_CENTER_OF_MASS)
        {
            /* TaskMap::Update(q, phi, J, H) base class (task_map.cpp:99-108): hessian(i) = J_row^T J_row */
            for (r = 0; r < m->length_jacobian; ++r)
                for (i = 0; i < N; ++i)
                    for (j = 0; j < N; ++j)
                        hm[(size_t)r * NN + (size_t)i * N + j] = jac[(size_t)i * lj + m->start_jacobian + r] * jac[(size_t)j * lj + m->start_jacobian + r];
            continue;
        }
        for (f = 0; f < m->n_frames; ++f)
        {
            const kframe_t* fr = &o->fr[m->frame_start + f];
            const double* kj = o->jac + (size_t)(m->frame_start + f) * 6 * N;
            compute_h(kj, N, fh);
            switch (m->kind)
            {
                case EXO_ORC_MAP_EFF_FRAME: /* eff_frame.cpp:75-90 */
                    memcpy(hm + (size_t)(6 * f) * NN, fh, 6 * NN * sizeof(double));
                    break;
                case EXO_ORC_MAP_EFF_POSITION: /* eff_position.cpp:57-70 */
                    memcpy(hm + (size_t)(3 * f) * NN, fh, 3 * NN * sizeof(double));
                    break;
                case EXO_ORC_MAP_EFF_POSITION_XY: /* eff_position_xy.cpp:59-75 */
                    memcpy(hm + (size_t)(2 * f) * NN, fh, 2 * NN * sizeof(double));
                    break;
                case EXO_ORC_MAP_EFF_ORIENTATION: /* eff_orientation.cpp:73-88 */
                    memcpy(hm + (size_t)(3 * f) * NN, fh + 3 * NN, 3 * NN * sizeof(double));
                    break;
                case EXO_ORC_MAP_DISTANCE: /* distance.cpp:59-91 */
                {
                    const double* p = fr->temp_ab.p;
                    const double nrm = vec_norm_kdl(p);
                    const double den = pow(p[0] * p[0] + p[1] * p[1] + p[2] * p[2], (double)(3 / 2)); /* integer division, as written */
                    for (i = 0; i < N; ++i)
                        for (j = 0; j < N; ++j)
                        {
                            const double ni = p[0] * kj[6 * i] + p[1] * kj[6 * i + 1] + p[2] * kj[6 * i + 2];
                            const double nj = p[0] * kj[6 * j] + p[1] * kj[6 * j + 1] + p[2] * kj[6 * j + 2];
                            const double part1 = -(ni * nj) / den;
                            const double part2 = ((p[0] * fh[(size_t)i * N + j] + p[1] * fh[NN + (size_t)i * N + j] + p[2] * fh[2 * NN + (size_t)i * N + j]) +
                                                  (kj[6 * i] * kj[6 * j] + kj[6 * i + 1] * kj[6 * j + 1] + kj[6 * i + 2] * kj[6 * j + 2])) / nrm;
                            hm[(size_t)f * NN + (size_t)i * N + j] = part1 + part2;
                        }
                    break;
                }
                default:
                    break;
            }
        }
    }
    free(phi);
    free(jac);
    free(fh);
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* IKSolver::Solve                                                      */
/* ------------------------------------------------------------------------------------------- */
/* Restates exotations/solvers/exotica_ik_solver/src/ik_solver.cpp:73-240 on an UnconstrainedEndPoseProblem whose task maps are
 * the oracle's maps (rho[n_maps]: S = diag(rho) per Jacobian row; rows of maps with rho 0 do not contribute).  q: start state
 * in, solution out.  Returns the TerminationCriterion value (planning_problem.h:49-62). */
static double ik_cost_at(exo_oracle* o, const double* q, const double* rho, const double* goal, double* phi, double* jac, double* yd)
{
    double c = 0.0;
    int k, i;
    exo_oracle_problem_update(o, q, phi, jac);
    exo_oracle_problem_ydiff(o, phi, goal, yd);
    for (k = 0; k < o->n_maps; ++k)
    {
        const taskmap_t* m = &o->maps[k];
        for (i = 0; i < m->length_jacobian; ++i) c += yd[m->start_jacobian + i] * rho[k] * yd[m->start_jacobian + i];
    }
    return c;
}

int exo_oracle_ik_solve(exo_oracle* o, double* q, const double* rho, const double* goal, const double* W, int max_iterations, double tolerance,
                        double lambda, double regmax, double gradient_tolerance, double step_tolerance, double* cost_out, int* iterations_out)
{
    /* init/ik_solver.in defaults of the parameters that only matter when MaxIterations == 1 */
    return exo_oracle_ik_solve_ex(o, q, rho, goal, W, max_iterations, tolerance, lambda, regmax, gradient_tolerance, step_tolerance, 0.0, 0.1, 0.5,
                                  1e-12, cost_out, iterations_out, NULL);
}

/* + MaxStep, ThresholdRegularizationIncrease / Decrease, MinimumRegularization (ik_solver.cpp:139-143,187-199; ik_solver.h:81-97);
 * lambda_out: lambda_ when Solve returns. */
int exo_oracle_ik_solve_ex(exo_oracle* o, double* q, const double* rho, const double* goal, const double* W, int max_iterations, double tolerance,
                           double lambda, double regmax, double gradient_tolerance, double step_tolerance, double max_step, double th_stepinc,
                           double th_stepdec, double regmin, double* cost_out, int* iterations_out, double* lambda_out)
{
    double steplength = 0.0; /* steplength_ (a member the reference never initialises; 0 here) */
    const int N = o->n_ctrl, lp = o->length_phi, lj = o->length_jac;
    double* phi = (double*)malloc((size_t)(lp + 1) * sizeof(double));
    double* jac = (double*)malloc((size_t)(lj + 1) * N * sizeof(double));
    double* yd = (double*)malloc((size_t)(lj + 1) * sizeof(double));
    double* cj = (double*)malloc((size_t)(lj + 1) * N * sizeof(double)); /* row-major lj x N */
    double* A = (double*)malloc((size_t)(lj + 1) * (lj + 1) * sizeof(double));
    double* L = (double*)malloc((size_t)(lj + 1) * (lj + 1) * sizeof(double));
    double* z = (double*)malloc((size_t)(lj + 1) * sizeof(double));
    double* qd = (double*)malloc((size_t)N * sizeof(double));
    double* qt = (double*)malloc((size_t)N * sizeof(double));
    double* srow = (double*)malloc((size_t)(lj + 1) * sizeof(double));
    int term = 0 /* IterationLimit */, it, k, i, j, r, c, ai;
    double error = 0.0, error_prev;
    for (k = 0; k < o->n_maps; ++k)
        for (i = 0; i < o->maps[k].length_jacobian; ++i) srow[o->maps[k].start_jacobian + i] = rho[k];
    for (it = 0; it < max_iterations; ++it)
    {
        double stop2 = 0.0, step = 0.0;
        int ok = 0, diverged = 0;
        error = ik_cost_at(o, q, rho, goal, phi, jac, yd);
        error_prev = error;
        if (error < tolerance)
        {
            term = 3; /* FunctionTolerance */
            break;
        }
        for (r = 0; r < lj; ++r)
        {
            yd[r] *= srow[r];
            for (c = 0; c < N; ++c)
            {
                cj[(size_t)r * N + c] = srow[r] * jac[(size_t)c * lj + r]; /* jac is column-major lj x N */
                stop2 += cj[(size_t)r * N + c] * cj[(size_t)r * N + c];
            }
        }
        for (r = 0; r < lj; ++r)
            for (j = 0; j < lj; ++j)
            {
                double acc = 0.0;
                for (c = 0; c < N; ++c) acc += cj[(size_t)r * N + c] * (W ? 1.0 / W[c] : 1.0) * cj[(size_t)j * N + c];
                A[(size_t)r * lj + j] = acc;
            }
        while (!ok)
        {
            ok = 1;
            for (j = 0; j < lj && ok; ++j)
            {
                double d = A[(size_t)j * lj + j] + lambda;
                for (k = 0; k < j; ++k) d -= L[(size_t)j * lj + k] * L[(size_t)j * lj + k];
                if (!(d > 0.0))
                {
                    ok = 0;
                    break;
                }
                L[(size_t)j * lj + j] = sqrt(d);
                for (i = j + 1; i < lj; ++i)
                {
                    double v = A[(size_t)i * lj + j];
                    for (k = 0; k < j; ++k) v -= L[(size_t)i * lj + k] * L[(size_t)j * lj + k];
                    L[(size_t)i * lj + j] = v / L[(size_t)j * lj + j];
                }
            }
            if (!ok)
            {
                lambda *= 10.0;
                if (lambda > regmax)
                {
                    lambda = regmax;
                    diverged = 1;
                    break;
                }
            }
        }
        if (diverged)
        {
            term = 5; /* Divergence */
            break;
        }
        for (i = 0; i < lj; ++i)
        {
            double v = yd[i];
            for (k = 0; k < i; ++k) v -= L[(size_t)i * lj + k] * z[k];
            z[i] = v / L[(size_t)i * lj + i];
        }
        for (i = lj - 1; i >= 0; --i)
        {
            double v = z[i];
            for (k = i + 1; k < lj; ++k) v -= L[(size_t)k * lj + i] * z[k];
            z[i] = v / L[(size_t)i * lj + i];
        }
        for (c = 0; c < N; ++c)
        {
            double acc = 0.0;
            for (r = 0; r < lj; ++r) acc += cj[(size_t)r * N + c] * z[r];
            qd[c] = (W ? 1.0 / W[c] : 1.0) * acc;
        }
        if (max_iterations == 1 && max_step != 0.0)
        {
            /* ScaleToStepSize (ik_solver.cpp:235-243), then q_ -= qd_: no line search */
            double max_vel = 0.0;
            for (c = 0; c < N; ++c) max_vel = fabs(qd[c]) > max_vel ? fabs(qd[c]) : max_vel;
            if (max_vel > max_step)
                for (c = 0; c < N; ++c) qd[c] = qd[c] * max_step / max_vel;
            for (c = 0; c < N; ++c) q[c] -= qd[c];
        }
        else
        for (ai = 0; ai < 10; ++ai)
        {
            /* Eigen::VectorXd::LinSpaced(10, 1.0, 0.1) (ik_solver.cpp:56); Eigen 3.3 (external) evaluates a descending LinSpaced
               from the high end: (i == 0) ? low : high - (size - 1 - i) * step */
            const double alpha = ai == 0 ? 1.0 : 0.1 - (9 - ai) * ((0.1 - 1.0) / 9.0);
            steplength = alpha;
            for (c = 0; c < N; ++c) qt[c] = q[c] - alpha * qd[c];
            error = ik_cost_at(o, qt, rho, goal, phi, jac, z /* scratch */);
            if (error < error_prev)
            {
                for (c = 0; c < N; ++c)
                {
                    q[c] = qt[c];
                    qd[c] *= alpha;
                }
                break;
            }
        }
        for (c = 0; c < N; ++c) step += qd[c] * qd[c];
        if (step < step_tolerance)
        {
            term = 2; /* StepTolerance */
            ++it;
            break;
        }
        if (sqrt(stop2) < gradient_tolerance)
        {
            term = 4; /* GradientTolerance */
            ++it;
            break;
        }
        /* adapt the regularisation to the step length (ik_solver.cpp:187-199) */
        if (max_iterations == 1 && max_step == 0.0 && steplength > th_stepdec)
        {
            lambda /= 10.0;
            if (lambda < regmin) lambda = regmin;
        }
        if (max_iterations == 1 && max_step == 0.0 && steplength <= th_stepinc)
        {
            lambda *= 10.0;
            if (lambda > regmax) lambda = regmax;
            if (lambda == regmax)
            {
                term = 5; /* Divergence */
                ++it;
                break;
            }
        }
    }
    if (lambda_out) *lambda_out = lambda;
    if (cost_out) *cost_out = ik_cost_at(o, q, rho, goal, phi, jac, yd);
    if (iterations_out) *iterations_out = it;
    free(phi); free(jac); free(yd); free(cj); free(A); free(L); free(z); free(qd); free(qt); free(srow);
    return term;
}
