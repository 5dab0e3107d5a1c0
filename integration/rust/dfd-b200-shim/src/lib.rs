//! `dfd-b200-shim`: binds `libdfd_b200.so` into datafusion-distributed without touching its planner.
//!
//! * [`GpuRepartitionExec`] — an `ExecutionPlan` with the schema, child and output partitioning of the
//!   `RepartitionExec(Partitioning::Hash)` it replaces; `execute(partition)` streams that destination's batches out of the
//!   GPU operator (`dfd_repartition_exec_*`).
//! * [`install_gpu_repartition_hook`] — `Worker::add_on_plan_hook` (src/worker/worker_service.rs:91-96, applied at
//!   src/worker/impl_set_plan.rs:122-124): every stage-head hash `RepartitionExec` whose keys are plain columns and whose
//!   schema and keys `dfd_repartition_supported` accepts is swapped one-for-one (no node added or removed, as the hook's contract asks).
//!
//! * [`GpuRepartitionCodec`] — `PhysicalExtensionCodec` for the node, for deployments that place it on the coordinator
//!   instead of through the worker hook.
//!
//! Not compiled in the build image (no Rust toolchain there); see Cargo.toml.
pub mod codec;
pub mod exec;
pub mod ffi;
pub mod hook;

pub use codec::GpuRepartitionCodec;
pub use exec::{GpuContext, GpuRepartitionExec, GpuRepartitionOptions};
pub use hook::{install_gpu_repartition_hook, rewrite_hash_repartitions};
