// rs_devmodel.h -- host-side conversion rs_model_desc (doubles) -> rs::DevModel (floats + derived tables).
#pragma once
#include <math.h>
#include <string.h>
#include "rs_core.cuh"

namespace rs {

inline void quat_to_mat_d(double* R, const double* q) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

// returns 0 on success, else a static error string via *err
inline int build_dev_model(const rs_model_desc* m, DevModel* D, const char** err) {
    memset(D, 0, sizeof *D);
    if (m->n_links < 0 || m->n_links > ML) { *err = "n_links out of range"; return 1; }
    if (m->n_geoms < 0 || m->n_pairs < 0 || m->n_act < 0 || m->n_feet < 0 || m->n_parts < 0 || m->n_dof < 0 || m->n_reset < 0 ||
        m->n_geoms > MG || m->n_pairs > MP || m->n_act > MA || m->n_feet > MF || m->n_parts > MPT) { *err = "model exceeds capacity"; return 1; }
    // a descriptor arrives through the C ABI from any caller: every index it carries is checked against its table
    if (m->n_dof > m->n_links) { *err = "n_dof exceeds n_links"; return 1; }
    if (m->n_reset > MA) { *err = "n_reset exceeds capacity"; return 1; }
    if (m->obs_dim < 0 || m->obs_dim > 8 + 2 * MA + MF) { *err = "obs_dim out of range"; return 1; }
    if (m->body_link < -1 || m->body_link >= m->n_links) { *err = "body_link out of range"; return 1; }
    if (m->task_link < -1 || m->task_link >= m->n_links || m->task_link2 < -1 || m->task_link2 >= m->n_links) { *err = "task link out of range"; return 1; }
    {
        int seen = 0;
        for (int i = 0; i < m->n_links; i++) {
            if (m->parent[i] < -1 || m->parent[i] >= i) { *err = "parent index out of order"; return 1; }
            if (m->jtype[i] == RS_JOINT_FIXED) { if (m->dof_idx[i] != -1) { *err = "fixed joint with a dof"; return 1; } }
            else { if (m->dof_idx[i] != seen) { *err = "dof_idx not consecutive"; return 1; } seen++; }
        }
        if (seen != m->n_dof) { *err = "n_dof does not match the joint types"; return 1; }
    }
    for (int k = 0; k < m->n_act; k++) if (m->act_dof[k] < 0 || m->act_dof[k] >= m->n_dof) { *err = "act_dof out of range"; return 1; }
    for (int k = 0; k < m->n_reset; k++) if (m->reset_dof[k] < 0 || m->reset_dof[k] >= m->n_dof) { *err = "reset_dof out of range"; return 1; }
    for (int k = 0; k < m->n_feet; k++) if (m->foot_link[k] < -1 || m->foot_link[k] >= m->n_links) { *err = "foot_link out of range"; return 1; }
    for (int k = 0; k < m->n_parts; k++) if (m->part_link[k] < -1 || m->part_link[k] >= m->n_links) { *err = "part_link out of range"; return 1; }
    for (int g = 0; g < m->n_geoms; g++) if (m->g_link[g] < -1 || m->g_link[g] >= m->n_links) { *err = "g_link out of range"; return 1; }
    for (int k = 0; k < m->n_pairs; k++) if (m->pair_a[k] < 0 || m->pair_a[k] >= m->n_geoms || m->pair_b[k] < 0 || m->pair_b[k] >= m->n_geoms) { *err = "pair geom out of range"; return 1; }
    if (m->task_kind != RS_TASK_WALKER) {
        const int nt = m->task_kind == RS_TASK_PENDULUM ? 2 : (m->task_kind == RS_TASK_DOUBLE_PENDULUM ? 3 : 4);
        for (int k = 0; k < nt; k++) if (m->task_dof[k] < 0 || m->task_dof[k] >= m->n_dof) { *err = "task_dof out of range"; return 1; }
        if ((m->task_kind == RS_TASK_DOUBLE_PENDULUM || m->task_kind == RS_TASK_REACHER) && m->task_link < 0) { *err = "task_link missing"; return 1; }
        if (m->task_kind == RS_TASK_REACHER && m->task_link2 < 0) { *err = "task_link2 missing"; return 1; }
    }
    if (m->max_contacts < 0) { *err = "max_contacts negative"; return 1; }
    D->n_links = m->n_links; D->n_dof = m->n_dof; D->fixed_base = m->fixed_base; D->n_geoms = m->n_geoms;
    D->n_pairs = m->n_pairs; D->n_act = m->n_act; D->n_feet = m->n_feet; D->n_parts = m->n_parts;
    D->obs_dim = m->obs_dim; D->has_floor = m->has_floor; D->n_iter = m->n_iter; D->frame_skip = m->frame_skip;
    D->alive_rule = m->alive_rule; D->body_link = m->body_link; D->max_episode_steps = m->max_episode_steps;
    D->max_contacts = m->max_contacts;
    D->base_mass = (float)m->base_mass;
    for (int k = 0; k < 3; k++) D->base_inertia[k] = (float)m->base_inertia[k];
    for (int i = 0; i < m->n_links; i++) {
        D->parent[i] = m->parent[i]; D->jtype[i] = m->jtype[i]; D->dof_idx[i] = m->dof_idx[i]; D->has_limit[i] = m->has_limit[i];
        if (m->dof_idx[i] >= 0) D->dof_link[m->dof_idx[i]] = i;
        double R0[9];
        quat_to_mat_d(R0, m->zero_rot[i]);
        for (int k = 0; k < 9; k++) D->R0[i][k] = (float)R0[k];
        const double* ax = m->axis[i];
        const double* d = m->d_vec[i];
        double sw[3] = {0, 0, 0}, sv[3] = {0, 0, 0};
        if (m->jtype[i] == RS_JOINT_REVOLUTE) {
            sw[0] = ax[0]; sw[1] = ax[1]; sw[2] = ax[2];
            sv[0] = ax[1] * d[2] - ax[2] * d[1]; sv[1] = ax[2] * d[0] - ax[0] * d[2]; sv[2] = ax[0] * d[1] - ax[1] * d[0];
        } else if (m->jtype[i] == RS_JOINT_PRISMATIC) { sv[0] = ax[0]; sv[1] = ax[1]; sv[2] = ax[2]; }
        for (int k = 0; k < 3; k++) {
            D->Sw[i][k] = (float)sw[k]; D->Sv[i][k] = (float)sv[k];
            D->axis[i][k] = (float)ax[k]; D->e_vec[i][k] = (float)m->e_vec[i][k]; D->d_vec[i][k] = (float)d[k];
            D->inertia[i][k] = (float)m->inertia[i][k];
        }
        D->mass[i] = (float)m->mass[i];
        D->lim_lo[i] = (float)m->lim_lo[i]; D->lim_hi[i] = (float)m->lim_hi[i];
        D->jdamping[i] = (float)m->jdamping[i]; D->max_vel[i] = (float)m->max_vel[i];
        D->armature[i] = (float)m->armature[i]; D->stiffness[i] = (float)m->stiffness[i];
    }
    for (int b = 0; b <= m->n_links; b++) D->break_thresh[b] = (float)m->break_thresh[b];
    for (int g = 0; g < m->n_geoms; g++) {
        D->g_link[g] = m->g_link[g]; D->g_type[g] = m->g_type[g]; D->g_floor[g] = m->g_floor[g];
        for (int k = 0; k < 3; k++) { D->g_p0[g][k] = (float)m->g_p0[g][k]; D->g_p1[g][k] = (float)m->g_p1[g][k]; }
        D->g_radius[g] = (float)m->g_radius[g];
        {
            double dx = m->g_p1[g][0] - m->g_p0[g][0], dy = m->g_p1[g][1] - m->g_p0[g][1], dz = m->g_p1[g][2] - m->g_p0[g][2];
            D->g_bound[g] = (float)(0.5 * sqrt(dx * dx + dy * dy + dz * dz) + m->g_radius[g]) * 1.0001f + 1e-6f;
        }
        for (int k = 0; k < 9; k++) D->g_rot[g][k] = (k % 4 == 0) ? 1.f : 0.f;
        if (m->g_type[g] == RS_GEOM_BOX) {
            // centre in both end slots (the segment helpers then see a point), bound = half diagonal, rotation from g_quat
            const double x = m->g_quat[g][0], y = m->g_quat[g][1], z = m->g_quat[g][2], w = m->g_quat[g][3];
            const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                                 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                                 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
            for (int k = 0; k < 9; k++) D->g_rot[g][k] = (float)R[k];
            double hx = m->g_p1[g][0], hy = m->g_p1[g][1], hz = m->g_p1[g][2];
            D->g_bound[g] = (float)sqrt(hx * hx + hy * hy + hz * hz) * 1.0001f + 1e-6f;
        }
        D->g_fric_floor[g] = (float)(m->g_friction[g] * m->floor_friction);
    }
    for (int k = 0; k < m->n_pairs; k++) {
        int a = m->pair_a[k], b = m->pair_b[k];
        D->pair_a[k] = (unsigned char)a; D->pair_b[k] = (unsigned char)b;
        D->pair_ss[k] = (m->g_type[a] == RS_GEOM_SPHERE && m->g_type[b] == RS_GEOM_SPHERE) ? 1 : 0;
        double ta = m->break_thresh[m->g_link[a] + 1], tb = m->break_thresh[m->g_link[b] + 1];
        D->pair_thresh[k] = (float)(ta < tb ? ta : tb);
        D->pair_friction[k] = (float)(m->g_friction[a] * m->g_friction[b]);
    }
    D->gravity = (float)m->gravity; D->timestep = (float)m->timestep;
    D->lin_damping = (float)m->lin_damping; D->ang_damping = (float)m->ang_damping;
    D->max_coord_vel = (float)m->max_coord_vel; D->erp = (float)m->erp; D->erp2 = (float)m->erp2;
    D->c_erp = (float)m->c_erp; D->c_erp2 = (float)m->c_erp2; D->c_split_thresh = (float)(m->c_split_thresh < -1e30 ? -1e30 : m->c_split_thresh);
    D->split_thresh = (float)m->split_thresh; D->max_limit_impulse = (float)m->max_limit_impulse; D->use_gyro = (float)m->use_gyro;
    for (int k = 0; k < m->n_act; k++) { D->act_dof[k] = m->act_dof[k]; D->act_gain[k] = (float)m->act_gain[k]; }
    for (int k = 0; k < m->n_feet; k++) D->foot_link[k] = m->foot_link[k];
    for (int k = 0; k < m->n_parts; k++) D->part_link[k] = m->part_link[k];
    D->alive_z = (float)m->alive_z; D->alive_bonus = (float)m->alive_bonus;
    D->electricity_cost = (float)m->electricity_cost; D->stall_torque_cost = (float)m->stall_torque_cost;
    D->joints_at_limit_cost = (float)m->joints_at_limit_cost; D->foot_collision_cost = (float)m->foot_collision_cost;
    D->initial_z = (float)m->initial_z;
    D->walk_target_x = m->walk_target_x; D->walk_target_y = m->walk_target_y; D->dt_env = m->dt_env;
    D->reset_joint_spread = (float)m->reset_joint_spread; D->reset_yaw_spread = (float)m->reset_yaw_spread;
    for (int k = 0; k < 3; k++) D->reset_base_pos[k] = (float)m->reset_base_pos[k];
    for (int k = 0; k < 4; k++) D->reset_base_quat[k] = (float)m->reset_base_quat[k];
    D->reset_set_base = m->reset_set_base;
    D->task_kind = m->task_kind; D->task_swingup = m->task_swingup; D->task_link = m->task_link; D->task_link2 = m->task_link2; D->n_reset = m->n_reset;
    for (int k = 0; k < m->n_reset && k < MA; k++) { D->reset_dof[k] = m->reset_dof[k]; D->reset_center[k] = (float)m->reset_center[k]; D->reset_spread[k] = (float)m->reset_spread[k]; }
    for (int k = 0; k < 4; k++) D->task_dof[k] = m->task_dof[k];
    D->flag_task = m->flag_task; D->flag_timeout_steps = m->flag_timeout_steps;
    D->flag_halflen = m->flag_halflen; D->flag_halfwidth = m->flag_halfwidth;
    D->has_cube = m->has_cube ? 1 : 0;
    if (m->has_cube) {
        if (!(m->cube_mass > 0) || !(m->cube_inertia[0] > 0) || !(m->cube_inertia[1] > 0) || !(m->cube_inertia[2] > 0) ||
            !(m->cube_half[0] > 0) || !(m->cube_half[1] > 0) || !(m->cube_half[2] > 0)) { *err = "cube mass / inertia / half extents must be positive"; return 1; }
        if (m->n_links + 1 > 254) { *err = "cube body index out of range"; return 1; }
        double b2 = 0;
        for (int k = 0; k < 3; k++) {
            D->cube_half[k] = (float)m->cube_half[k]; D->cube_inertia[k] = (float)m->cube_inertia[k]; D->cube_start[k] = (float)m->cube_start[k];
            b2 += m->cube_half[k] * m->cube_half[k];
        }
        D->cube_mass = (float)m->cube_mass; D->cube_thresh = (float)m->cube_break_thresh;
        D->cube_bound = (float)sqrt(b2) * 1.0001f + 1e-6f;
        D->cube_fric_floor = (float)(m->cube_friction * m->floor_friction);
        for (int g = 0; g < m->n_geoms; g++) D->cube_fric_geom[g] = (float)(m->g_friction[g] * m->cube_friction);
    }
    return 0;
}

}  // namespace rs
