//! resolve2d_b200.zig — Zig binding that mirrors resolve2d's `Solver` + `EntityFactory` (src/core/lib.zig:13-315)
//! over the C ABI of libr2d_b200.so (include/r2d_abi.h).
//!
//! UNVERIFIED: there is no Zig toolchain in the build image, so this file has never been compiled.  It is the
//! reference-side stub a maintainer would add next to src/root.zig; see INTEGRATION.md.
//! Link with:  exe.linkSystemLibrary("r2d_b200"); exe.addLibraryPath(.{ .cwd_relative = "resolve2d_b200" });
const std = @import("std");

pub const Id = u32; // reference: u16 (Bodies/RigidBody.zig:15)

pub const Error = error{ OutOfMemory, InvalidRigidBodyId, NoSuchIdExists, InvalidArgument, NoDevice, Cuda, ColorOverflow, BadState, GridRange };

fn check(status: c_int) Error!void {
    return switch (status) {
        0 => {},
        -1 => error.OutOfMemory,
        -2 => error.InvalidRigidBodyId,
        -3 => error.NoSuchIdExists,
        -4 => error.InvalidArgument,
        -5 => error.NoDevice,
        -7 => error.ColorOverflow,
        -8 => error.BadState,
        -9 => error.GridRange,
        else => error.Cuda,
    };
}

pub const Vector2 = extern struct { x: f32 = 0, y: f32 = 0 };

const c = struct {
    pub const r2d_solver = opaque {};
    pub const BodyOpts = extern struct { pos_x: f32, pos_y: f32, vel_x: f32, vel_y: f32, angle: f32, omega: f32, mu: f32, mass_value: f32, mass_is_density: i32 };
    pub const JointParams = extern struct { power_max: f32, power_min: f32, beta: f32 };
    pub const BodyState = extern struct {
        id: u32, shape: i32, is_static: i32,
        pos_x: f32, pos_y: f32, angle: f32, momentum_x: f32, momentum_y: f32, ang_momentum: f32,
        force_x: f32, force_y: f32, torque: f32, mass: f32, inertia: f32, mu: f32,
        aabb_x: f32, aabb_y: f32, aabb_half_w: f32, aabb_half_h: f32, shape_a: f32, shape_b: f32,
    };
    pub extern "c" fn r2d_create(cell_width: f32, table_mult: u32, device: c_int, out: *?*r2d_solver) c_int;
    pub extern "c" fn r2d_destroy(s: ?*r2d_solver) c_int;
    pub extern "c" fn r2d_clear(s: ?*r2d_solver) c_int;
    pub extern "c" fn r2d_process(s: ?*r2d_solver, dt: f32, sub_steps: u32, collision_iters: u32) c_int;
    pub extern "c" fn r2d_process_read(s: ?*r2d_solver, dt: f32, sub_steps: u32, collision_iters: u32, ids: ?[*]u32, pos_xy: ?[*]f32, angle: ?[*]f32, momentum_xy: ?[*]f32, ang_momentum: ?[*]f32, aabb_xywh: ?[*]f32, n: usize) c_int;
    pub extern "c" fn r2d_make_disc(s: ?*r2d_solver, o: *const BodyOpts, radius: f32, out_id: *u32) c_int;
    pub extern "c" fn r2d_make_rect(s: ?*r2d_solver, o: *const BodyOpts, width: f32, height: f32, out_id: *u32) c_int;
    pub extern "c" fn r2d_make_gravity(s: ?*r2d_solver, g: f32) c_int;
    pub extern "c" fn r2d_make_distance_joint(s: ?*r2d_solver, p: *const JointParams, id1: u32, id2: u32, target: f32, out: *usize) c_int;
    pub extern "c" fn r2d_make_offset_distance_joint(s: ?*r2d_solver, p: *const JointParams, id1: u32, id2: u32, r1x: f32, r1y: f32, r2x: f32, r2y: f32, target: f32, out: *usize) c_int;
    pub extern "c" fn r2d_make_fixed_position_joint(s: ?*r2d_solver, p: *const JointParams, id: u32, tx: f32, ty: f32, out: *usize) c_int;
    pub extern "c" fn r2d_make_motor_joint(s: ?*r2d_solver, p: *const JointParams, id: u32, omega: f32, out: *usize) c_int;
    pub extern "c" fn r2d_exclude_pair(s: ?*r2d_solver, id1: u32, id2: u32) c_int;
    pub extern "c" fn r2d_remove_body(s: ?*r2d_solver, id: u32) c_int;
    pub extern "c" fn r2d_num_bodies(s: ?*r2d_solver, out: *usize) c_int;
    pub extern "c" fn r2d_body_id_at(s: ?*r2d_solver, i: usize, out: *u32) c_int;
    pub extern "c" fn r2d_body_get(s: ?*r2d_solver, id: u32, out: *BodyState) c_int;
    pub extern "c" fn r2d_body_set_static(s: ?*r2d_solver, id: u32, v: c_int) c_int;
    pub extern "c" fn r2d_body_set_momentum(s: ?*r2d_solver, id: u32, x: f32, y: f32) c_int;
    pub extern "c" fn r2d_body_set_ang_momentum(s: ?*r2d_solver, id: u32, l: f32) c_int;
    pub extern "c" fn r2d_body_set_force(s: ?*r2d_solver, id: u32, x: f32, y: f32) c_int;
    pub extern "c" fn r2d_body_set_torque(s: ?*r2d_solver, id: u32, t: f32) c_int;
    pub extern "c" fn r2d_set_mode(s: ?*r2d_solver, mode: c_int) c_int; // 0 parity, 1 fast, 2 reference order (validation)
    pub extern "c" fn r2d_set_option(s: ?*r2d_solver, option: c_int, value: u32) c_int; // 1 warm start, 2 sleeping, 3 sleep calls
    // a batch of worlds sharded over several GPUs (world w -> shard w * G / n_worlds), include/r2d_abi.h "r2d_sharded_*"
    pub const r2d_sharded = opaque {};
    pub extern "c" fn r2d_sharded_create(n_worlds: u32, devices: [*]const c_int, n_devices: u32, cell_width: f32, table_mult: u32, out: *?*r2d_sharded) c_int;
    pub extern "c" fn r2d_sharded_destroy(b: ?*r2d_sharded) c_int;
    pub extern "c" fn r2d_sharded_world(b: ?*r2d_sharded, world: u32, out: *?*r2d_solver) c_int;
    pub extern "c" fn r2d_sharded_process(b: ?*r2d_sharded, dt: f32, sub_steps: u32, collision_iters: u32) c_int;
    pub extern "c" fn r2d_sharded_process_read(b: ?*r2d_sharded, dt: f32, sub_steps: u32, collision_iters: u32, ids: ?[*]u32, pos_xy: ?[*]f32, angle: ?[*]f32, momentum_xy: ?[*]f32, ang_momentum: ?[*]f32, aabb_xywh: ?[*]f32, n: usize) c_int;
    pub extern "c" fn r2d_sharded_write_forces(b: ?*r2d_sharded, force_xy_torque: [*]const f32, n: usize) c_int;
    pub extern "c" fn r2d_read_bodies(s: ?*r2d_solver, ids: ?[*]u32, pos_xy: ?[*]f32, angle: ?[*]f32, momentum_xy: ?[*]f32, ang_momentum: ?[*]f32, aabb_xywh: ?[*]f32, capacity: usize) c_int;
};

pub const BodyState = c.BodyState;

/// Constraint.Parameters (Constraints/Constraint.zig:27-31)
pub const Parameters = struct {
    power_max: f32 = std.math.inf(f32),
    power_min: f32 = -std.math.inf(f32),
    beta: f32 = 10,
    fn toC(self: Parameters) c.JointParams {
        return .{ .power_max = self.power_max, .power_min = self.power_min, .beta = self.beta };
    }
};

pub const EntityFactory = struct {
    pub const BodyHandle = struct {
        id: Id,
        solver: *Solver,
        /// `body_unwrap().*` by value: the state lives in HBM, there is no `*RigidBody` to hand out (lib.zig:23-31).
        pub fn body(self: BodyHandle) ?BodyState {
            var st: BodyState = undefined;
            check(c.r2d_body_get(self.solver.handle, self.id, &st)) catch return null;
            return st;
        }
        pub fn setStatic(self: BodyHandle, v: bool) Error!void {
            try check(c.r2d_body_set_static(self.solver.handle, self.id, @intFromBool(v)));
        }
        pub fn setTorque(self: BodyHandle, t: f32) Error!void {
            try check(c.r2d_body_set_torque(self.solver.handle, self.id, t));
        }
        pub fn setForce(self: BodyHandle, f: Vector2) Error!void {
            try check(c.r2d_body_set_force(self.solver.handle, self.id, f.x, f.y));
        }
        pub fn setAngularMomentum(self: BodyHandle, l: f32) Error!void {
            try check(c.r2d_body_set_ang_momentum(self.solver.handle, self.id, l));
        }
    };
    pub const ConstraintHandle = usize;
    pub const BodyOptions = struct {
        pos: Vector2,
        vel: Vector2 = .{},
        angle: f32 = 0.0,
        omega: f32 = 0.0,
        mu: f32 = 0.5,
        mass_prop: union(enum) { density: f32, mass: f32 },
        fn toC(self: BodyOptions) c.BodyOpts {
            return .{
                .pos_x = self.pos.x, .pos_y = self.pos.y, .vel_x = self.vel.x, .vel_y = self.vel.y,
                .angle = self.angle, .omega = self.omega, .mu = self.mu,
                .mass_value = switch (self.mass_prop) { .density => |d| d, .mass => |m| m },
                .mass_is_density = switch (self.mass_prop) { .density => 1, .mass => 0 },
            };
        }
    };
    pub const DiscOptions = struct { radius: f32 = 1.0 };
    pub const RectangleOptions = struct { width: f32 = 1.0, height: f32 = 0.5 };

    solver: *Solver,
    const Self = @This();

    pub fn makeDiscBody(self: *Self, bo: BodyOptions, go: DiscOptions) Error!BodyHandle {
        var id: u32 = 0;
        const o = bo.toC();
        try check(c.r2d_make_disc(self.solver.handle, &o, go.radius, &id));
        return .{ .id = id, .solver = self.solver };
    }
    pub fn makeRectangleBody(self: *Self, bo: BodyOptions, go: RectangleOptions) Error!BodyHandle {
        var id: u32 = 0;
        const o = bo.toC();
        try check(c.r2d_make_rect(self.solver.handle, &o, go.width, go.height, &id));
        return .{ .id = id, .solver = self.solver };
    }
    pub fn makeDownwardsGravity(self: *Self, g: f32) Error!void {
        try check(c.r2d_make_gravity(self.solver.handle, g));
    }
    pub fn makeOffsetDistanceJoint(self: *Self, params: Parameters, h1: BodyHandle, h2: BodyHandle, r1: Vector2, r2: Vector2, target_distance: f32) Error!ConstraintHandle {
        var idx: usize = 0;
        const p = params.toC();
        try check(c.r2d_make_offset_distance_joint(self.solver.handle, &p, h1.id, h2.id, r1.x, r1.y, r2.x, r2.y, target_distance, &idx));
        return idx;
    }
    pub fn makeDistanceJoint(self: *Self, params: Parameters, h1: BodyHandle, h2: BodyHandle, target_distance: f32) Error!ConstraintHandle {
        var idx: usize = 0;
        const p = params.toC();
        try check(c.r2d_make_distance_joint(self.solver.handle, &p, h1.id, h2.id, target_distance, &idx));
        return idx;
    }
    pub fn makeFixedPositionJoint(self: *Self, params: Parameters, h: BodyHandle, target_position: Vector2) Error!ConstraintHandle {
        var idx: usize = 0;
        const p = params.toC();
        try check(c.r2d_make_fixed_position_joint(self.solver.handle, &p, h.id, target_position.x, target_position.y, &idx));
        return idx;
    }
    pub fn makeMotorJoint(self: *Self, params: Parameters, h: BodyHandle, target_omega: f32) Error!ConstraintHandle {
        var idx: usize = 0;
        const p = params.toC();
        try check(c.r2d_make_motor_joint(self.solver.handle, &p, h.id, target_omega, &idx));
        return idx;
    }
    pub fn excludeCollisionPair(self: *Self, h1: BodyHandle, h2: BodyHandle) Error!void {
        try check(c.r2d_exclude_pair(self.solver.handle, h1.id, h2.id));
    }
};

pub const Solver = struct {
    handle: ?*c.r2d_solver,
    const Self = @This();

    /// Solver.init (lib.zig:145); the allocator argument is gone — device memory is owned by the library.
    pub fn init(spatialhash_cell_width: f32, spatialhash_table_size_mult: usize) Error!Self {
        var h: ?*c.r2d_solver = null;
        try check(c.r2d_create(spatialhash_cell_width, @intCast(spatialhash_table_size_mult), 0, &h));
        return .{ .handle = h };
    }
    pub fn deinit(self: *Self) void {
        _ = c.r2d_destroy(self.handle);
        self.handle = null;
    }
    pub fn clear(self: *Self) Error!void {
        try check(c.r2d_clear(self.handle));
    }
    /// Solver.process (lib.zig:189)
    pub fn process(self: *Self, dt: f32, sub_steps: usize, collision_iters: usize) Error!void {
        try check(c.r2d_process(self.handle, dt, @intCast(sub_steps), @intCast(collision_iters)));
    }
    /// process() followed by a bulk read of every body (iteration order) in one call and one host synchronisation; any
    /// slice may be null.  Replaces "solver.process(...); for (solver.bodies.values()) |b| ..." of the demos' frame loops.
    pub fn processRead(self: *Self, dt: f32, sub_steps: usize, collision_iters: usize, ids: ?[]u32, pos_xy: ?[]f32, angle: ?[]f32, momentum_xy: ?[]f32, ang_momentum: ?[]f32, aabb_xywh: ?[]f32) Error!void {
        try check(c.r2d_process_read(self.handle, dt, @intCast(sub_steps), @intCast(collision_iters), if (ids) |s| s.ptr else null, if (pos_xy) |s| s.ptr else null, if (angle) |s| s.ptr else null, if (momentum_xy) |s| s.ptr else null, if (ang_momentum) |s| s.ptr else null, if (aabb_xywh) |s| s.ptr else null, self.numBodies()));
    }
    /// Roadmap items of the reference (README.md:59-64), off by default: warm starting, sleeping.
    pub const Option = enum(c_int) { warm_start = 1, sleeping = 2, sleep_calls = 3 };
    pub fn setOption(self: *Self, option: Option, value: u32) Error!void {
        try check(c.r2d_set_option(self.handle, @intFromEnum(option), value));
    }
    pub fn bodyHandle(self: *Self, id: Id) EntityFactory.BodyHandle {
        return .{ .id = id, .solver = self };
    }
    pub fn removeRigidBody(self: *Self, id: Id) Error!void {
        try check(c.r2d_remove_body(self.handle, id));
    }
    pub fn entityFactory(self: *Self) EntityFactory {
        return .{ .solver = self };
    }
    pub fn numBodies(self: *Self) usize {
        var n: usize = 0;
        _ = c.r2d_num_bodies(self.handle, &n);
        return n;
    }
};
