//! The worker-side plan hook: swap every eligible `RepartitionExec(Hash)` for [`GpuRepartitionExec`].

use std::sync::Arc;

use datafusion::common::tree_node::{Transformed, TreeNode};
use datafusion::physical_plan::repartition::RepartitionExec;
use datafusion::physical_plan::ExecutionPlan;
use datafusion_distributed::Worker;

use crate::exec::{GpuContext, GpuRepartitionExec, GpuRepartitionOptions};

/// One-for-one rewrite (the hook's contract: "the same, or equivalent in terms of execution" — no node is added or
/// removed, src/worker/worker_service.rs:86-90). Nodes the GPU operator does not serve are returned unchanged, and so is
/// the whole plan if the traversal fails.
pub fn rewrite_hash_repartitions(
    plan: Arc<dyn ExecutionPlan>,
    ctx: &Arc<GpuContext>,
    options: GpuRepartitionOptions,
) -> Arc<dyn ExecutionPlan> {
    let original = Arc::clone(&plan);
    plan.transform_down(|node| {
        if let Some(r) = node.as_any().downcast_ref::<RepartitionExec>() {
            if let Some(gpu) = GpuRepartitionExec::try_from_repartition(r, Arc::clone(ctx), options) {
                return Ok(Transformed::yes(Arc::new(gpu) as Arc<dyn ExecutionPlan>));
            }
        }
        Ok(Transformed::no(node))
    })
    .map(|t| t.data)
    .unwrap_or(original)
}

/// `examples/localhost_worker.rs`-style installation:
/// ```ignore
/// let mut worker = Worker::default();
/// install_gpu_repartition_hook(&mut worker, GpuContext::try_new(gpu_index)?, GpuRepartitionOptions::default());
/// Server::builder().add_service(worker.into_flight_server()).serve(addr).await?;
/// ```
pub fn install_gpu_repartition_hook(worker: &mut Worker, ctx: Arc<GpuContext>, options: GpuRepartitionOptions) {
    worker.add_on_plan_hook(move |plan| rewrite_hash_repartitions(plan, &ctx, options));
}
