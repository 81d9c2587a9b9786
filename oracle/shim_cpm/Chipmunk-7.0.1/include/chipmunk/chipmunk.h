// Interface stand-in for the slice of Chipmunk2D 7.0.1 (not vendored by the reference: eQsetup1.sh clones it)
// that src/abm/cpmEcoli.cpp and src/abm/Ecoli.cpp use, written here so that the reference's own cell classes
// can be compiled IN PLACE and run as a parity pin for the rod geometry (oracle/Makefile ->
// oracle/_ref/libeq_cell_ref.so).  TEST INFRASTRUCTURE ONLY.
//
// Real: rigid-body state and the transform arithmetic the point-in-rod predicate goes through
// (cpBodySetPosition / cpBodySetAngle -> body transform; cpBodyWorldToLocal = cpTransformPoint of
// cpTransformRigidInverse; cpvcross, cpv arithmetic), restated from Chipmunk 7.0.1's published cpBody.c /
// cpTransform.h / cpVect.h with the same operation order.  Inert: space, shapes, constraints, forces -- no
// physics step is ever taken here.
#pragma once
#include <cmath>
#include <cstdlib>

typedef double cpFloat;
typedef void *cpDataPointer;
typedef unsigned char cpBool;
#define cpTrue 1
#define cpFalse 0
struct cpVect { cpFloat x, y; };
struct cpTransform { cpFloat a, b, c, d, tx, ty; };
typedef enum { CP_BODY_TYPE_DYNAMIC, CP_BODY_TYPE_KINEMATIC, CP_BODY_TYPE_STATIC } cpBodyType;

struct cpSpace { int unused; };
struct cpShape { int unused; };
struct cpConstraint { int unused; };
struct cpPointQueryInfo { int unused; };
struct cpBody;
typedef void (*cpBodyVelocityFunc)(cpBody *body, cpVect gravity, cpFloat damping, cpFloat dt);
struct cpBody {
    cpFloat m, m_inv, i, i_inv;
    cpVect cog, p, v, f;
    cpFloat a, w, t;
    cpTransform transform;
    cpDataPointer userData;
    cpBodyType type;
    cpBodyVelocityFunc velocity_func;
};

static const cpVect cpvzero = {0.0, 0.0};
static inline cpVect cpv(const cpFloat x, const cpFloat y) { cpVect v = {x, y}; return v; }
static inline cpVect cpvadd(const cpVect a, const cpVect b) { return cpv(a.x + b.x, a.y + b.y); }
static inline cpVect cpvsub(const cpVect a, const cpVect b) { return cpv(a.x - b.x, a.y - b.y); }
static inline cpVect cpvneg(const cpVect v) { return cpv(-v.x, -v.y); }
static inline cpVect cpvmult(const cpVect v, const cpFloat s) { return cpv(v.x * s, v.y * s); }
static inline cpFloat cpvdot(const cpVect a, const cpVect b) { return a.x * b.x + a.y * b.y; }
static inline cpFloat cpvcross(const cpVect a, const cpVect b) { return a.x * b.y - a.y * b.x; }
static inline cpFloat cpvlength(const cpVect v) { return std::sqrt(cpvdot(v, v)); }
static inline cpFloat cpvdist(const cpVect a, const cpVect b) { return cpvlength(cpvsub(a, b)); }
static inline cpVect cpvnormalize(const cpVect v) { return cpvmult(v, 1.0 / (cpvlength(v) + 2.2250738585072014e-308)); }
static inline cpFloat cpvtoangle(const cpVect v) { return std::atan2(v.y, v.x); }
static inline cpVect cpvforangle(const cpFloat a) { return cpv(std::cos(a), std::sin(a)); }
// chipmunk.h's C++ operators
static inline cpVect operator*(const cpVect v, const cpFloat s) { return cpvmult(v, s); }
static inline cpVect operator+(const cpVect a, const cpVect b) { return cpvadd(a, b); }
static inline cpVect operator-(const cpVect a, const cpVect b) { return cpvsub(a, b); }
static inline cpVect operator-(const cpVect v) { return cpvneg(v); }

// cpTransform.h
static inline cpTransform cpTransformNewTranspose(cpFloat a, cpFloat c, cpFloat tx, cpFloat b, cpFloat d, cpFloat ty)
{
    cpTransform t = {a, b, c, d, tx, ty};
    return t;
}
static inline cpTransform cpTransformRigidInverse(cpTransform t)
{
    return cpTransformNewTranspose(t.d, -t.c, (t.c * t.ty - t.tx * t.d), -t.b, t.a, (t.tx * t.b - t.a * t.ty));
}
static inline cpVect cpTransformPoint(cpTransform t, cpVect p) { return cpv(t.a * p.x + t.c * p.y + t.tx, t.b * p.x + t.d * p.y + t.ty); }

// cpBody.c: SetTransform(body, p, a) with the centre of gravity at the origin
static inline void cp_shim_set_transform(cpBody *body)
{
    const cpVect rot = cpvforangle(body->a);
    const cpVect c = body->cog;
    body->transform = cpTransformNewTranspose(rot.x, -rot.y, body->p.x - (c.x * rot.x - c.y * rot.y),
                                              rot.y, rot.x, body->p.y - (c.x * rot.y + c.y * rot.x));
}
static inline cpBody *cpBodyNew(cpFloat mass, cpFloat moment)
{
    cpBody *b = (cpBody *)std::calloc(1, sizeof(cpBody));
    b->m = mass; b->m_inv = mass != 0.0 ? 1.0 / mass : 0.0;
    b->i = moment; b->i_inv = moment != 0.0 ? 1.0 / moment : 0.0;
    b->type = CP_BODY_TYPE_DYNAMIC;
    cp_shim_set_transform(b);
    return b;
}
static inline void cpBodyFree(cpBody *b) { std::free(b); }
// cpBody.c: p = TransformVect(transform, cog) + position; the centre of gravity is the origin here
static inline void cpBodySetPosition(cpBody *b, cpVect p) { b->p = p; cp_shim_set_transform(b); }
static inline void cpBodySetAngle(cpBody *b, cpFloat a) { b->a = a; cp_shim_set_transform(b); }
static inline void cpBodySetVelocity(cpBody *b, cpVect v) { b->v = v; }
static inline cpVect cpBodyGetPosition(const cpBody *b) { return cpTransformPoint(b->transform, cpvzero); }
// cpBody.c (7.0.1): the unit rotation vector stored by cpBodySetAngle, (cos a, sin a)
static inline cpVect cpBodyGetRotation(const cpBody *b) { return cpv(b->transform.a, b->transform.b); }
static inline cpFloat cpBodyGetAngle(const cpBody *b) { return b->a; }
static inline cpVect cpBodyGetVelocity(const cpBody *b) { return b->v; }
static inline cpFloat cpBodyGetAngularVelocity(const cpBody *b) { return b->w; }
static inline cpBodyType cpBodyGetType(cpBody *b) { return b->type; }
static inline void cpBodySetType(cpBody *b, cpBodyType t) { b->type = t; }
static inline void cpBodySetUserData(cpBody *b, cpDataPointer d) { b->userData = d; }
static inline cpDataPointer cpBodyGetUserData(const cpBody *b) { return b->userData; }
static inline void cpBodySetVelocityUpdateFunc(cpBody *b, cpBodyVelocityFunc f) { b->velocity_func = f; }
static inline cpVect cpBodyLocalToWorld(const cpBody *b, const cpVect p) { return cpTransformPoint(b->transform, p); }
static inline cpVect cpBodyWorldToLocal(const cpBody *b, const cpVect p) { return cpTransformPoint(cpTransformRigidInverse(b->transform), p); }
static inline void cpBodyApplyForceAtWorldPoint(cpBody *b, cpVect f, cpVect) { b->f = cpvadd(b->f, f); }
static inline void cpBodyApplyForceAtLocalPoint(cpBody *b, cpVect f, cpVect) { b->f = cpvadd(b->f, f); }

// inert: space / shapes / constraints
static inline cpSpace *cpSpaceNew() { return (cpSpace *)std::calloc(1, sizeof(cpSpace)); }
static inline void cpSpaceFree(cpSpace *s) { std::free(s); }
static inline void cpSpaceSetCollisionSlop(cpSpace *, cpFloat) {}
static inline void cpSpaceSetCollisionBias(cpSpace *, cpFloat) {}
static inline void cpSpaceUseSpatialHash(cpSpace *, cpFloat, int) {}
static inline void cpSpaceSetIterations(cpSpace *, int) {}
static inline void cpSpaceSetDamping(cpSpace *, cpFloat) {}
static inline void cpSpaceStep(cpSpace *, cpFloat) {}
static inline cpBody *cpSpaceGetStaticBody(cpSpace *) { static cpBody b; return &b; }
static inline cpBody *cpSpaceAddBody(cpSpace *, cpBody *b) { return b; }
static inline void cpSpaceRemoveBody(cpSpace *, cpBody *) {}
static inline cpShape *cpSpaceAddShape(cpSpace *, cpShape *s) { return s; }
static inline void cpSpaceRemoveShape(cpSpace *, cpShape *) {}
static inline cpConstraint *cpSpaceAddConstraint(cpSpace *, cpConstraint *c) { return c; }
static inline void cpSpaceRemoveConstraint(cpSpace *, cpConstraint *) {}
static inline cpShape *cp_shim_shape() { return (cpShape *)std::calloc(1, sizeof(cpShape)); }
static inline cpConstraint *cp_shim_constraint() { return (cpConstraint *)std::calloc(1, sizeof(cpConstraint)); }
static inline cpShape *cpSegmentShapeNew(cpBody *, cpVect, cpVect, cpFloat) { return cp_shim_shape(); }
static inline cpShape *cpBoxShapeNew(cpBody *, cpFloat, cpFloat, cpFloat) { return cp_shim_shape(); }
static inline cpShape *cpCircleShapeNew(cpBody *, cpFloat, cpVect) { return cp_shim_shape(); }
static inline cpShape *cpPolyShapeNew(cpBody *, int, const cpVect *, cpTransform, cpFloat) { return cp_shim_shape(); }
static inline void cpShapeFree(cpShape *s) { std::free(s); }
static inline void cpShapeSetFriction(cpShape *, cpFloat) {}
static inline void cpShapeSetElasticity(cpShape *, cpFloat) {}
static inline void cpShapeCacheBB(cpShape *) {}
static inline cpFloat cpShapePointQuery(const cpShape *, cpVect, cpPointQueryInfo *) { return 0.0; }
static inline void cpPolyShapeSetVertsRaw(cpShape *, int, cpVect *) {}
static inline void cpCircleShapeSetOffset(cpShape *, cpVect) {}
static inline cpConstraint *cpGrooveJointNew(cpBody *, cpBody *, cpVect, cpVect, cpVect) { return cp_shim_constraint(); }
static inline cpConstraint *cpDampedSpringNew(cpBody *, cpBody *, cpVect, cpVect, cpFloat, cpFloat, cpFloat) { return cp_shim_constraint(); }
static inline void cpGrooveJointSetGrooveA(cpConstraint *, cpVect) {}
static inline void cpGrooveJointSetGrooveB(cpConstraint *, cpVect) {}
static inline void cpGrooveJointSetAnchorB(cpConstraint *, cpVect) {}
static inline void cpDampedSpringSetRestLength(cpConstraint *, cpFloat) {}
static inline cpFloat cpDampedSpringGetRestLength(const cpConstraint *) { return 0.0; }
static inline void cpConstraintSetMaxForce(cpConstraint *, cpFloat) {}
static inline void cpConstraintSetErrorBias(cpConstraint *, cpFloat) {}
static inline void cpConstraintSetMaxBias(cpConstraint *, cpFloat) {}
static inline void cpConstraintSetCollideBodies(cpConstraint *, cpBool) {}
static inline void cpConstraintFree(cpConstraint *c) { std::free(c); }
#define cpAssertSoft(...)
#define cpAssertSaneBody(body)
