//! `PhysicalExtensionCodec` for [`GpuRepartitionExec`] — only needed when the node is placed BEFORE the stage plan is
//! shipped (a physical optimizer rule on the coordinator); with the worker-side plan hook (`hook.rs`) the node is created
//! after decoding and never crosses the wire. Registered like any user codec of the reference
//! (`with_distributed_user_codec`, src/distributed_planner/session_builder.rs; example: examples/custom_execution_plan.rs:246-292).
//!
//! Wire format (little endian, no protobuf dependency): magic "DFDG" | version u32 = 1 | num_partitions u32 |
//! chunk_rows i64 | max_pinned_chunks i32 | n_keys u32 | key column indices i32[n_keys]. The child plan travels as the
//! node's single input, as for every `ExecutionPlan`.

use std::sync::Arc;

use datafusion::error::{DataFusionError, Result};
use datafusion::execution::TaskContext;
use datafusion::physical_expr::expressions::Column;
use datafusion::physical_expr::PhysicalExpr;
use datafusion::physical_plan::repartition::RepartitionExec;
use datafusion::physical_plan::{ExecutionPlan, Partitioning};
use datafusion_proto::physical_plan::PhysicalExtensionCodec;

use crate::exec::{GpuContext, GpuRepartitionExec, GpuRepartitionOptions};

const MAGIC: &[u8; 4] = b"DFDG";

#[derive(Debug)]
pub struct GpuRepartitionCodec {
    /// the decoding worker's GPU (one `dfd_ctx` per GPU per process)
    pub ctx: Arc<GpuContext>,
}

fn bad(msg: &str) -> DataFusionError {
    DataFusionError::Internal(format!("GpuRepartitionCodec: {msg}"))
}

struct Reader<'a>(&'a [u8]);
impl Reader<'_> {
    fn take<const N: usize>(&mut self) -> Result<[u8; N]> {
        if self.0.len() < N {
            return Err(bad("truncated buffer"));
        }
        let (head, rest) = self.0.split_at(N);
        self.0 = rest;
        Ok(head.try_into().unwrap())
    }
}

impl PhysicalExtensionCodec for GpuRepartitionCodec {
    fn try_decode(&self, buf: &[u8], inputs: &[Arc<dyn ExecutionPlan>], _ctx: &TaskContext) -> Result<Arc<dyn ExecutionPlan>> {
        let [input] = inputs else {
            return Err(bad("GpuRepartitionExec has exactly one child"));
        };
        let mut r = Reader(buf);
        if &r.take::<4>()? != MAGIC {
            return Err(bad("not a GpuRepartitionExec node"));
        }
        if u32::from_le_bytes(r.take()?) != 1 {
            return Err(bad("unknown version"));
        }
        let num_partitions = u32::from_le_bytes(r.take()?) as usize;
        let options = GpuRepartitionOptions {
            chunk_rows: i64::from_le_bytes(r.take()?),
            max_pinned_chunks: i32::from_le_bytes(r.take()?),
        };
        let n_keys = u32::from_le_bytes(r.take()?) as usize;
        let schema = input.schema();
        let mut exprs: Vec<Arc<dyn PhysicalExpr>> = Vec::with_capacity(n_keys);
        for _ in 0..n_keys {
            let idx = i32::from_le_bytes(r.take()?) as usize;
            let field = schema.fields().get(idx).ok_or_else(|| bad("key column out of range"))?;
            exprs.push(Arc::new(Column::new(field.name(), idx)));
        }
        // schema, equivalence properties and advertised partitioning are exactly those of the CPU operator it stands for
        let cpu = RepartitionExec::try_new(Arc::clone(input), Partitioning::Hash(exprs, num_partitions))?;
        GpuRepartitionExec::try_from_repartition(&cpu, Arc::clone(&self.ctx), options)
            .map(|g| Arc::new(g) as Arc<dyn ExecutionPlan>)
            .ok_or_else(|| DataFusionError::NotImplemented("schema not supported by the GPU operator on this worker".into()))
    }

    fn try_encode(&self, node: Arc<dyn ExecutionPlan>, buf: &mut Vec<u8>) -> Result<()> {
        let Some(g) = node.as_any().downcast_ref::<GpuRepartitionExec>() else {
            return Err(bad(&format!("expected GpuRepartitionExec, got {}", node.name())));
        };
        buf.extend_from_slice(MAGIC);
        buf.extend_from_slice(&1u32.to_le_bytes());
        buf.extend_from_slice(&(g.num_partitions() as u32).to_le_bytes());
        buf.extend_from_slice(&g.options().chunk_rows.to_le_bytes());
        buf.extend_from_slice(&g.options().max_pinned_chunks.to_le_bytes());
        buf.extend_from_slice(&(g.key_columns().len() as u32).to_le_bytes());
        for k in g.key_columns() {
            buf.extend_from_slice(&k.to_le_bytes());
        }
        Ok(())
    }
}
