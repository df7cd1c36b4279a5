// compiler/src/cuda_lowering.rs  (new; `mod cuda_lowering;` in compiler/src/lib.rs, `WriteCuda` re-exported from
// translating_traits/mod.rs beside WriteC / WriteWasm, mod.rs:5-29)
//
// produce_c (template.rs:281-472 + the bucket emitters) prints C++ that RE-EXECUTES circom's compile-time control flow
// at run time: loops over `lvar` counters, index arithmetic, sub-component bookkeeping.  The CUDA producer executes
// that part HERE, once, by partial evaluation of the IR: every variable slot is either Known(BigInt) - evaluated with
// circom_algebra::modular_arithmetic, exactly what the C++ would compute - or Dynamic(Ref), a value that depends on
// signals and therefore becomes a .cb2c operation.  Template bodies come out as straight-line op lists; function bodies
// (whose control flow may depend on run-time values) keep their loops as register code with jumps.
//
// Written against circom 2.2.3, not compiled in circom_b200's build image.  `Err(())` marks what format version 1
// does not express (docs/CB2C.md, last section).
use crate::circuit_design::{function::FunctionCodeInfo, template::TemplateCodeInfo};
use crate::intermediate_representation::ir_interface::*;
use circom_algebra::modular_arithmetic as ma;
use code_producers::components::IODef;
use code_producers::cuda_elements::*;
use num_bigint_dig::BigInt;

#[derive(Clone)]
pub enum Val { Known(BigInt), Dynamic(Ref) }

pub struct TemplateCtx<'a> {
    pub producer: &'a CUDAProducer,
    pub file: &'a mut Cb2cFile,
    pub q: BigInt,
    pub rec: TemplateRecord,
    pub vars: Vec<Val>,                 // lvar[]: template.rs:288
    pub sub_of_cmp: Vec<u32>,           // component slot -> index into rec.subs (CreateCmpBucket order)
    pub n_out: usize, pub n_in: usize,
    pub signal_extents: &'a [(usize, usize)],   // (first local id, total size) of every declared signal (array) of the template
}

pub trait WriteCuda {
    /// evaluates the instruction; a value-producing instruction returns where its value is
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()>;
}

impl<'a> TemplateCtx<'a> {
    fn tmp(&mut self) -> Ref { self.rec.n_tmp += 1; Ref::Tmp(self.rec.n_tmp - 1) }
    fn emit(&mut self, op: Op, a: Ref, b: Ref, c: Ref) -> Ref {
        let d = self.tmp();
        self.rec.ops.push(OpRec { op, d, a, b, c });
        d
    }
    fn as_ref(&mut self, v: &Val) -> Ref {
        match v {
            Val::Known(k) => { let q = self.q.clone(); Ref::Const(self.file.const_id(k, &q)) }
            Val::Dynamic(r) => *r,
        }
    }
    /// `Fr_toInt` of an address expression: addresses inside templates are compile-time values
    fn address(&mut self, i: &InstructionPointer) -> Result<usize, ()> {
        match i.produce_cuda(self)? {
            Some(Val::Known(k)) => k.to_u64_digits().1.first().map(|x| *x as usize).or(Some(0)).ok_or(()),
            _ => Err(()), // run-time addressing (ToAddress of a signal-dependent value): not in version 1
        }
    }
    /// signal index inside this template -> reference (outputs, inputs, intermediates: executed_template.rs:442-552)
    fn own(&self, idx: usize) -> Ref { Ref::Own(idx as u32) }
    /// number of `stride`-sized elements from own signal `base` to the end of the array that contains it
    fn signal_extent(&self, base: usize, stride: usize) -> Result<usize, ()> {
        let (first, len) = self.signal_extents.iter().find(|(f, l)| *f <= base && base < f + l).ok_or(())?;
        if stride == 0 || (first + len - base) % stride != 0 { return Err(()); }
        Ok((first + len - base) / stride)
    }
    /// A load address that depends on a signal (`out <-- table[sel]`): the location is address arithmetic over one
    /// `ToAddress(value)` (compute_bucket.rs:361-363: Fr_toInt) - base + toInt(value) * stride.  Returns (base, value, stride).
    fn dynamic_address(&mut self, i: &InstructionPointer) -> Result<(usize, Ref, usize), ()> {
        if let Instruction::Compute(c) = &**i {
            use OperatorType::*;
            match c.op {
                ToAddress => {
                    let v = c.stack[0].produce_cuda(self)?.ok_or(())?;
                    return Ok((0, self.as_ref(&v), 1));
                }
                AddAddress => {
                    let (k, d) = match (self.address(&c.stack[0]), self.address(&c.stack[1])) {
                        (Ok(k), Err(())) => (k, &c.stack[1]),
                        (Err(()), Ok(k)) => (k, &c.stack[0]),
                        _ => return Err(()),                              // two run-time parts: not expanded
                    };
                    let (base, v, stride) = self.dynamic_address(d)?;
                    return Ok((base + k, v, stride));
                }
                MulAddress => {
                    let (k, d) = match (self.address(&c.stack[0]), self.address(&c.stack[1])) {
                        (Ok(k), Err(())) => (k, &c.stack[1]),
                        (Err(()), Ok(k)) => (k, &c.stack[0]),
                        _ => return Err(()),
                    };
                    let (base, v, stride) = self.dynamic_address(d)?;
                    return Ok((base * k, v, stride * k));
                }
                _ => {}
            }
        }
        Err(())
    }
    /// The expansion of such a load over the array it indexes (docs/CB2C.md, "run-time addresses"; the Python DSL's
    /// Template.expanded_ops writes the same ops): EQ(value, i), SELECT(eq_i ? signal[base + i * stride] : 0), two adder trees,
    /// ASSERT(sum of the eq_i).  `extent` = elements from `base` to the end of the indexed array:
    /// TemplateCodeInfo::signal_extents (added by circom/src/patch.md from TemplateInstance::wires, the source of
    /// build_input_output_list, build.rs:498-520).
    fn indexed_signal_load(&mut self, base: usize, value: Ref, stride: usize, extent: usize) -> Ref {
        let q = self.q.clone();
        let zero = Ref::Const(self.file.const_id(&BigInt::from(0), &q));
        let mut eqs = vec![];
        let mut sels = vec![];
        for i in 0..extent {
            let ci = Ref::Const(self.file.const_id(&BigInt::from(i), &q));
            let e = self.emit(Op::EQ, value, ci, Ref::None);
            eqs.push(e);
            sels.push(self.emit(Op::SELECT, self.own(base + i * stride), zero, e));
        }
        let tree = |cx: &mut Self, mut xs: Vec<Ref>| -> Ref {
            while xs.len() > 1 {
                let mut nxt = vec![];
                for p in xs.chunks(2) { nxt.push(if p.len() == 2 { cx.emit(Op::ADD, p[0], p[1], Ref::None) } else { p[0] }); }
                xs = nxt;
            }
            xs[0]
        };
        let hit = tree(self, eqs);
        self.rec.ops.push(OpRec { op: Op::ASSERT, d: Ref::None, a: hit, b: Ref::None, c: Ref::None });
        tree(self, sels)
    }
    /// `LocationRule::Mapped` (store_bucket.rs:498-566, load_bucket.rs:264-322): a signal of an element of a component
    /// array of mixed templates.  The C++ looks the offset up at run time in templateInsId2IOSignalInfo[templateId of the
    /// element]; here the element (cmp) is a compile-time value, so its template instance is known and the very same
    /// arithmetic runs now: defs[signal_code].offset + ((i0 * lengths[1] + i1) * lengths[2] + ...) * size, missing trailing
    /// indexes multiplied through.  Bus fields (AccessType::Qualified -> busInsId2FieldInfo) are flattened the same way
    /// from producer.busid_field_info.
    fn mapped_address(&mut self, cmp: usize, signal_code: usize, indexes: &[AccessType]) -> Result<usize, ()> {
        let template_id = self.rec.subs[self.sub_of_cmp[cmp] as usize] as usize;
        let defs = self.producer.io_map.get(&template_id).ok_or(())?;
        let mut cur = defs.iter().find(|d| d.code == signal_code).ok_or(())?.clone();
        let mut offset = cur.offset;
        for (pos, access) in indexes.iter().enumerate() {
            match access {
                AccessType::Indexed(info) => {
                    let mut idx = self.address(&info.indexes[0])?;
                    for i in 1..info.indexes.len() { idx = idx * cur.lengths[i] + self.address(&info.indexes[i])?; }
                    if info.indexes.len() < info.symbol_dim {
                        if pos + 1 != indexes.len() { return Err(()); }             // must be the last access (load_bucket.rs:297)
                        for i in info.indexes.len()..info.symbol_dim { idx *= cur.lengths[i]; }
                    }
                    offset += idx * cur.size;
                }
                AccessType::Qualified(field_no) => {
                    let bus = cur.bus_id.ok_or(())?;
                    let f = self.producer.busid_field_info.get(bus).and_then(|fs| fs.get(*field_no)).ok_or(())?;
                    offset += f.offset;
                    cur = IODef { code: *field_no, offset: f.offset, lengths: f.dimensions.clone(), size: f.size, bus_id: f.bus_id };
                }
            }
        }
        Ok(offset)
    }
}

fn operator_code(op: &OperatorType) -> Result<Op, ()> {
    use OperatorType::*;
    Ok(match op {
        Mul => Op::MUL, Div => Op::DIV, Add => Op::ADD, Sub => Op::SUB, Pow => Op::POW, IntDiv => Op::IDIV, Mod => Op::MOD,
        ShiftL => Op::SHL, ShiftR => Op::SHR, LesserEq => Op::LEQ, GreaterEq => Op::GEQ, Lesser => Op::LT, Greater => Op::GT,
        Eq(SizeOption::Single(1)) => Op::EQ, NotEq => Op::NEQ, BoolOr => Op::LOR, BoolAnd => Op::LAND, BoolNot => Op::LNOT,
        BitOr => Op::BOR, BitAnd => Op::BAND, BitXor => Op::BXOR, Complement => Op::BNOT, PrefixSub => Op::NEG,
        Eq(_) => return Err(()),               // array equality: expand element-wise upstream
        ToAddress | MulAddress | AddAddress => unreachable!("address arithmetic is evaluated, never emitted"),
    })
}

fn eval_known(op: &OperatorType, a: &BigInt, b: Option<&BigInt>, q: &BigInt) -> Result<BigInt, ()> {
    use OperatorType::*;
    let b0 = BigInt::from(0);
    let b = b.unwrap_or(&b0);
    Ok(match op {
        Mul | MulAddress => ma::mul(a, b, q), Add | AddAddress => ma::add(a, b, q), Sub => ma::sub(a, b, q),
        Div => ma::div(a, b, q).map_err(|_| {})?, IntDiv => ma::idiv(a, b, q).map_err(|_| {})?,
        Mod => ma::mod_op(a, b, q).map_err(|_| {})?, Pow => ma::pow(a, b, q),
        ShiftL => ma::shift_l(a, b, q).map_err(|_| {})?, ShiftR => ma::shift_r(a, b, q).map_err(|_| {})?,
        LesserEq => ma::lesser_eq(a, b, q), GreaterEq => ma::greater_eq(a, b, q), Lesser => ma::lesser(a, b, q),
        Greater => ma::greater(a, b, q), Eq(_) => ma::eq(a, b, q), NotEq => ma::not_eq(a, b, q),
        BoolOr => ma::bool_or(a, b, q), BoolAnd => ma::bool_and(a, b, q), BoolNot => ma::not(a, q),
        BitOr => ma::bit_or(a, b, q), BitAnd => ma::bit_and(a, b, q), BitXor => ma::bit_xor(a, b, q),
        Complement => ma::complement_256(a, q), PrefixSub => ma::prefix_sub(a, q), ToAddress => a.clone(),
    })
}

impl WriteCuda for Instruction {
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        use Instruction::*;
        match self {
            Value(v) => v.produce_cuda(cx), Load(v) => v.produce_cuda(cx), Store(v) => v.produce_cuda(cx),
            Compute(v) => v.produce_cuda(cx), Call(v) => v.produce_cuda(cx), Branch(v) => v.produce_cuda(cx),
            Return(_) => Err(()),                 // only inside functions (FunctionCtx below)
            Assert(v) => v.produce_cuda(cx), Log(v) => v.produce_cuda(cx),
            Loop(v) => v.produce_cuda(cx), CreateCmp(v) => v.produce_cuda(cx),
        }
    }
}

impl WriteCuda for ValueBucket {   // value_bucket.rs:81-87
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        Ok(Some(Val::Known(match self.parse_as {
            ValueType::U32 => BigInt::from(self.value),
            ValueType::BigInt => cx.producer.field_tracking[self.value].parse::<BigInt>().map_err(|_| {})?,
        })))
    }
}

impl WriteCuda for ComputeBucket { // compute_bucket.rs:276-400
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        let mut args = vec![];
        for a in &self.stack { args.push(a.produce_cuda(cx)?.ok_or(())?); }
        if let (Some(Val::Known(a)), b) = (args.get(0), args.get(1)) {
            let kb = match b { Some(Val::Known(k)) => Some(Some(k)), None => Some(None), _ => None };
            if let Some(kb) = kb { let q = cx.q.clone(); return Ok(Some(Val::Known(eval_known(&self.op, a, kb, &q)?))); }
        }
        let ra = cx.as_ref(&args[0]);
        let rb = if args.len() > 1 { cx.as_ref(&args[1]) } else { Ref::None };
        Ok(Some(Val::Dynamic(cx.emit(operator_code(&self.op)?, ra, rb, Ref::None))))
    }
}

impl WriteCuda for LoadBucket {    // load_bucket.rs:325-447
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        let idx = match (&self.src, &self.address_type) {
            (LocationRule::Indexed { location, .. }, AddressType::Signal) => match cx.address(location) {
                Ok(k) => k,
                Err(()) => {                           // a signal array indexed by a signal: expanded over the array
                    let (base, value, stride) = cx.dynamic_address(location)?;
                    let extent = cx.signal_extent(base, stride)?;
                    return Ok(Some(Val::Dynamic(cx.indexed_signal_load(base, value, stride, extent))));
                }
            },
            (LocationRule::Indexed { location, .. }, _) => cx.address(location)?,
            (LocationRule::Mapped { signal_code, indexes }, AddressType::SubcmpSignal { cmp_address, .. }) => {
                let cmp = cx.address(cmp_address)?;
                cx.mapped_address(cmp, *signal_code, indexes)?
            }
            _ => return Err(()),                       // Mapped is only ever produced for sub-component signals
        };
        Ok(Some(match &self.address_type {
            AddressType::Variable => cx.vars[idx].clone(),
            AddressType::Signal => Val::Dynamic(cx.own(idx)),
            AddressType::SubcmpSignal { cmp_address, .. } => {
                let cmp = cx.address(cmp_address)?;
                Val::Dynamic(Ref::Sub { sub: cx.sub_of_cmp[cmp], idx: idx as u32 })
            }
        }))
    }
}

impl WriteCuda for StoreBucket {   // store_bucket.rs:607-834
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        if self.context.size != SizeOption::Single(1) { return Err(()); } // array copies: expanded by the caller per element
        let v = self.src.produce_cuda(cx)?.ok_or(())?;
        let idx = match (&self.dest, &self.dest_address_type) {
            (LocationRule::Indexed { location, .. }, _) => cx.address(location)?,
            (LocationRule::Mapped { signal_code, indexes }, AddressType::SubcmpSignal { cmp_address, .. }) => {
                let cmp = cx.address(cmp_address)?;
                cx.mapped_address(cmp, *signal_code, indexes)?
            }
            _ => return Err(()),
        };
        match &self.dest_address_type {
            AddressType::Variable => { cx.vars[idx] = v; }
            AddressType::Signal => { let a = cx.as_ref(&v); let d = cx.own(idx); cx.rec.ops.push(OpRec { op: Op::COPY, d, a, b: Ref::None, c: Ref::None }); }
            AddressType::SubcmpSignal { cmp_address, input_information, .. } => {
                // the trigger (inputCounter, StatusInput::Last / NoLast, store_bucket.rs:660-734) is re-derived by the
                // lowering from the order of the stores; a trigger that depends on run-time values is not expressible
                // StatusInput::Last / NoLast / Unknown (store_bucket.rs:663-734) need no distinction here: the lowering counts
                // the stores into a component's inputs in execution order and runs it at the last one - the run-time rule of
                // `Unknown`, of which `Last` / `NoLast` are the compiler's pre-computed cases
                let _ = input_information;
                let cmp = cx.address(cmp_address)?;
                let a = cx.as_ref(&v);
                cx.rec.ops.push(OpRec { op: Op::COPY, d: Ref::Sub { sub: cx.sub_of_cmp[cmp], idx: idx as u32 }, a, b: Ref::None, c: Ref::None });
            }
        }
        Ok(None)
    }
}

impl WriteCuda for LogBucket {     // log_bucket.rs:104-162: one LOG op per argument; the library prints from the witness afterwards
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        let n = self.argsprint.len();
        for (k, arg) in self.argsprint.iter().enumerate() {
            let last = Ref::Imm(if k + 1 == n { 1 } else { 0 });
            match arg {
                LogBucketArg::LogStr(id) => {
                    let text = cx.producer.string_table[*id].clone();
                    if text.is_empty() || text.bytes().any(|b| b < 0x20 || b >= 0x7f || b == b'%' || b == b'\\' || b == b'"') { return Err(()); }
                    let sid = cx.file.string_id(&text);
                    cx.rec.ops.push(OpRec { op: Op::LOG, d: Ref::None, a: Ref::None, b: Ref::Imm(sid), c: last });
                }
                LogBucketArg::LogExp(e) => {
                    let a = match e.produce_cuda(cx)?.ok_or(())? {
                        Val::Known(k) => { let q = cx.q.clone(); Ref::Const(cx.file.const_id(&k, &q)) }
                        Val::Dynamic(r @ (Ref::Own(_) | Ref::Sub { .. } | Ref::One)) => r,
                        Val::Dynamic(_) => return Err(()),   // an expression no signal holds: not printable from the witness
                    };
                    cx.rec.ops.push(OpRec { op: Op::LOG, d: Ref::None, a, b: Ref::None, c: last });
                }
            }
        }
        Ok(None)
    }
}

impl WriteCuda for LoopBucket {    // loop_bucket.rs:76-91: unrolled - the condition must be compile-time
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        loop {
            match self.continue_condition.produce_cuda(cx)? {
                Some(Val::Known(c)) => { if c == BigInt::from(0) { return Ok(None); } }
                _ => return Err(()), // a template loop on a signal-dependent condition cannot generate constraints either
            }
            for i in &self.body { i.produce_cuda(cx)?; }
        }
    }
}

impl WriteCuda for BranchBucket {  // branch_bucket.rs:100-122
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        match self.cond.produce_cuda(cx)? {
            Some(Val::Known(c)) => {
                for i in if c != BigInt::from(0) { &self.if_branch } else { &self.else_branch } { i.produce_cuda(cx)?; }
                Ok(None)
            }
            Some(Val::Dynamic(c)) => {
                // `var v = cond ? a : b` / if-else over variables with a run-time condition: both arms are evaluated on
                // copies of the variable state and merged with SELECT (total: x/0 = 0).  Arms that store signals or
                // create components under a run-time condition are refused.
                let before = cx.vars.clone();
                let n_ops = cx.rec.ops.len();
                for i in &self.if_branch { i.produce_cuda(cx)?; }
                let then_vars = std::mem::replace(&mut cx.vars, before);
                for i in &self.else_branch { i.produce_cuda(cx)?; }
                if cx.rec.ops[n_ops..].iter().any(|o| !matches!(o.d, Ref::Tmp(_))) { return Err(()); }
                for k in 0..cx.vars.len() {
                    let (t, e) = (then_vars[k].clone(), cx.vars[k].clone());
                    let same = match (&t, &e) { (Val::Known(x), Val::Known(y)) => x == y, (Val::Dynamic(x), Val::Dynamic(y)) => x == y, _ => false };
                    if !same {
                        let (rt, re) = (cx.as_ref(&t), cx.as_ref(&e));
                        cx.vars[k] = Val::Dynamic(cx.emit(Op::SELECT, rt, re, c));
                    }
                }
                Ok(None)
            }
            None => Err(()),
        }
    }
}

impl WriteCuda for AssertBucket {  // assert_bucket.rs:70-88
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        if self.is_constraint_equality && cx.producer.sanity_check_style == 0 { return Ok(None); }
        // `a === b` arrives as Compute(Eq, [a, b]): keep the two sides so that the lowering can compare representations
        if let Instruction::Compute(c) = self.evaluate.as_ref() {
            if let OperatorType::Eq(SizeOption::Single(1)) = c.op {
                let a = c.stack[0].produce_cuda(cx)?.ok_or(())?;
                let b = c.stack[1].produce_cuda(cx)?.ok_or(())?;
                let (ra, rb) = (cx.as_ref(&a), cx.as_ref(&b));
                cx.rec.ops.push(OpRec { op: Op::ASSERT_EQ, d: Ref::None, a: ra, b: rb, c: Ref::None });
                return Ok(None);
            }
        }
        let v = self.evaluate.produce_cuda(cx)?.ok_or(())?;
        let a = cx.as_ref(&v);
        cx.rec.ops.push(OpRec { op: Op::ASSERT, d: Ref::None, a, b: Ref::None, c: Ref::None });
        Ok(None)
    }
}

impl WriteCuda for CreateCmpBucket { // create_component_bucket.rs:204-352
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        if self.is_part_mixed_array_not_uniform_parallel { return Err(()); }   // mixed-template arrays: templateInsId2IOSignalInfo
        let first = cx.address(&self.sub_cmp_id)?;
        for (k, _parallel) in &self.defined_positions {            // `parallel` is ignored: every component is data-parallel here
            let slot = first + *k;
            if cx.sub_of_cmp.len() <= slot { cx.sub_of_cmp.resize(slot + 1, u32::MAX); }
            cx.sub_of_cmp[slot] = cx.rec.subs.len() as u32;
            cx.rec.subs.push(self.template_id as u32);             // template ids = order of Circuit::templates
            // symbols: `name_subcomponent` plus the indices of position k inside `dimensions` (row-major), e.g. `bits[2]`
            let mut name = self.name_subcomponent.clone();
            let mut rest = *k;
            let mut idx = vec![0usize; self.dimensions.len()];
            for (d, dim) in self.dimensions.iter().enumerate().rev() { idx[d] = rest % dim; rest /= dim; }
            for i in idx { name.push_str(&format!("[{}]", i)); }
            cx.rec.sub_names.push(name);
        }
        Ok(None)
    }
}

impl WriteCuda for CallBucket {    // call_bucket.rs:466-533
    fn produce_cuda(&self, cx: &mut TemplateCtx) -> Result<Option<Val>, ()> {
        if self.argument_types.iter().any(|t| t.size != SizeOption::Single(1)) { return Err(()); } // array arguments: one ARG per element upstream
        let fid = *cx.producer.function_ids.get(&self.symbol).ok_or(())?;
        for a in &self.arguments {
            let v = a.produce_cuda(cx)?.ok_or(())?;
            let r = cx.as_ref(&v);
            cx.rec.ops.push(OpRec { op: Op::ARG, d: Ref::None, a: r, b: Ref::None, c: Ref::None });
        }
        // `var r[n] = f(..)`: ONE call, its n results in n consecutive temporaries (CALL operand c = n, docs/CB2C.md)
        let n_res = match &self.return_info {
            ReturnType::Final(f) => match f.context.size { SizeOption::Single(n) => n, _ => return Err(()) },
            _ => 1,
        };
        if n_res == 0 || n_res > 64 { return Err(()); }
        let d = cx.tmp();
        let d0 = match d { Ref::Tmp(i) => i, _ => return Err(()) };
        for _ in 1..n_res { cx.tmp(); }
        cx.rec.ops.push(OpRec { op: Op::CALL, d, a: Ref::Imm(fid), b: Ref::Imm(self.arguments.len() as u32),
                                c: if n_res > 1 { Ref::Imm(n_res as u32) } else { Ref::None } });
        match &self.return_info {
            ReturnType::Intermediate { .. } => Ok(Some(Val::Dynamic(d))),
            ReturnType::Final(f) if n_res > 1 => {                  // array destination: variables only (signals: per-element COPY)
                let idx = match &f.dest { LocationRule::Indexed { location, .. } => cx.address(location)?, _ => return Err(()) };
                for k in 0..n_res {
                    let r = Ref::Tmp(d0 + k as u32);
                    match &f.dest_address_type {
                        AddressType::Variable => cx.vars[idx + k] = Val::Dynamic(r),
                        AddressType::Signal => { let o = cx.own(idx + k); cx.rec.ops.push(OpRec { op: Op::COPY, d: o, a: r, b: Ref::None, c: Ref::None }); }
                        AddressType::SubcmpSignal { cmp_address, .. } => {
                            let cmp = cx.address(cmp_address)?;
                            cx.rec.ops.push(OpRec { op: Op::COPY, d: Ref::Sub { sub: cx.sub_of_cmp[cmp], idx: (idx + k) as u32 }, a: r, b: Ref::None, c: Ref::None });
                        }
                    }
                }
                Ok(None)
            }
            ReturnType::Final(f) => {                               // `x <-- f(...)`: the store is part of the bucket
                let idx = match &f.dest { LocationRule::Indexed { location, .. } => cx.address(location)?, _ => return Err(()) };
                match &f.dest_address_type {
                    AddressType::Variable => cx.vars[idx] = Val::Dynamic(d),
                    AddressType::Signal => { let o = cx.own(idx); cx.rec.ops.push(OpRec { op: Op::COPY, d: o, a: d, b: Ref::None, c: Ref::None }); }
                    AddressType::SubcmpSignal { cmp_address, .. } => {
                        let cmp = cx.address(cmp_address)?;
                        cx.rec.ops.push(OpRec { op: Op::COPY, d: Ref::Sub { sub: cx.sub_of_cmp[cmp], idx: idx as u32 }, a: d, b: Ref::None, c: Ref::None });
                    }
                }
                Ok(None)
            }
        }
    }
}

// Function bodies keep their control flow: registers = the function's variable slots (FunctionCodeInfo::
// max_number_of_vars) + expression temporaries; LoopBucket -> JZ / JMP, BranchBucket -> JZ / JMP, ReturnBucket -> RET,
// Load/Store of `Variable` with a run-time index -> LOADX / STOREX (compute_bucket.rs:361-363 ToAddress = Fr_toInt).
// The walk is the same as above with `Val::Dynamic(Ref::Tmp(reg))` everywhere; see docs/CB2C.md "function".
// One pass over a function body.  `regs` = variable slots first (LoadBucket / StoreBucket with AddressType::Variable
// address them; a compile-time address is the register itself, a run-time one goes through LOADX / STOREX with base 0),
// expression temporaries after them.
struct FunctionCtx<'a> { producer: &'a CUDAProducer, file: &'a mut Cb2cFile, q: BigInt, rec: &'a mut FunctionRecord }
impl<'a> FunctionCtx<'a> {
    fn reg(&mut self) -> Ref { self.rec.n_regs += 1; Ref::Tmp(self.rec.n_regs - 1) }
    fn push(&mut self, op: Op, d: Ref, a: Ref, b: Ref, c: Ref) -> usize { self.rec.code.push(OpRec { op, d, a, b, c }); self.rec.code.len() - 1 }
    fn value(&mut self, i: &InstructionPointer) -> Result<Ref, ()> {
        Ok(match i.as_ref() {
            Instruction::Value(v) => {
                let k = match v.parse_as { ValueType::U32 => BigInt::from(v.value),
                                           ValueType::BigInt => self.producer.field_tracking[v.value].parse::<BigInt>().map_err(|_| {})? };
                let q = self.q.clone();
                Ref::Const(self.file.const_id(&k, &q))
            }
            Instruction::Compute(c) => {
                // address arithmetic (ToAddress / MulAddress / AddAddress) is ordinary arithmetic on registers here
                let a = self.value(&c.stack[0])?;
                let b = if c.stack.len() > 1 { self.value(&c.stack[1])? } else { Ref::None };
                let op = match c.op { OperatorType::ToAddress => return Ok(a), OperatorType::MulAddress => Op::MUL,
                                      OperatorType::AddAddress => Op::ADD, _ => operator_code(&c.op)? };
                let d = self.reg();
                self.push(op, d, a, b, Ref::None);
                d
            }
            Instruction::Load(l) => match (&l.address_type, &l.src) {
                (AddressType::Variable, LocationRule::Indexed { location, .. }) => match location.as_ref() {
                    Instruction::Value(v) => Ref::Tmp(v.value as u32),
                    _ => { let idx = self.value(location)?; let d = self.reg(); self.push(Op::LOADX, d, Ref::Imm(0), idx, Ref::None); d }
                },
                _ => return Err(()),   // functions only see variables (call_bucket.rs: arguments are copied in)
            },
            Instruction::Call(c) => {
                // a call inside a function body (call_bucket.rs:466-533): the arguments go to consecutive fresh registers
                // (the callee's parameters; array arguments element by element), CALL {d, function, first argument register,
                // result count}.  The callee must already be lowered: `produce_cb2c` walks the functions callee first.
                let fid = *self.producer.function_ids.get(&c.symbol).ok_or(())?;
                if fid >= self.file.functions.len() as u32 { return Err(()); }      // recursion / not yet lowered
                let mut vals = Vec::new();
                for a in &c.arguments { vals.push(self.value(a)?); }
                let base = self.rec.n_regs;
                for v in vals { let r = self.reg(); self.push(Op::COPY, r, v, Ref::None, Ref::None); }
                let n_res = match &c.return_info {
                    ReturnType::Final(f) => match f.context.size { SizeOption::Single(n) => n, _ => return Err(()) },
                    _ => 1,
                };
                if n_res == 0 || n_res > 64 { return Err(()); }
                let d = self.reg();
                for _ in 1..n_res { self.reg(); }
                self.push(Op::CALL, d, Ref::Imm(fid), if c.arguments.is_empty() { Ref::None } else { Ref::Tmp(base) },
                          if n_res > 1 { Ref::Imm(n_res as u32) } else { Ref::None });
                if let ReturnType::Final(f) = &c.return_info {      // `x = g(..)` / `var r[n] = g(..)`: store into the variable slots
                    let d0 = match d { Ref::Tmp(i) => i, _ => return Err(()) };
                    if let (AddressType::Variable, LocationRule::Indexed { location, .. }) = (&f.dest_address_type, &f.dest) {
                        if let Instruction::Value(v) = location.as_ref() {
                            for k in 0..n_res { self.push(Op::COPY, Ref::Tmp(v.value as u32 + k as u32), Ref::Tmp(d0 + k as u32), Ref::None, Ref::None); }
                        } else { return Err(()); }
                    } else { return Err(()); }
                }
                d
            }
            _ => return Err(()),
        })
    }
    fn stmt(&mut self, i: &InstructionPointer) -> Result<(), ()> {
        match i.as_ref() {
            Instruction::Store(s) => {
                let v = self.value(&s.src)?;
                match (&s.dest_address_type, &s.dest) {
                    (AddressType::Variable, LocationRule::Indexed { location, .. }) => match location.as_ref() {
                        Instruction::Value(a) => { self.push(Op::COPY, Ref::Tmp(a.value as u32), v, Ref::None, Ref::None); }
                        _ => { let idx = self.value(location)?; self.push(Op::STOREX, Ref::None, Ref::Imm(0), idx, v); }
                    },
                    _ => return Err(()),
                }
            }
            Instruction::Loop(l) => {                      // loop_bucket.rs:76-91: while (Fr_isTrue(cond)) body
                let head = self.rec.code.len();
                let c = self.value(&l.continue_condition)?;
                let jz = self.push(Op::JZ, Ref::None, c, Ref::Imm(0), Ref::None);
                for b in &l.body { self.stmt(b)?; }
                self.push(Op::JMP, Ref::None, Ref::Imm(head as u32), Ref::None, Ref::None);
                let end = self.rec.code.len() as u32;
                self.rec.code[jz].b = Ref::Imm(end);
            }
            Instruction::Branch(b) => {                    // branch_bucket.rs:100-122
                let c = self.value(&b.cond)?;
                let jz = self.push(Op::JZ, Ref::None, c, Ref::Imm(0), Ref::None);
                for x in &b.if_branch { self.stmt(x)?; }
                let jmp = self.push(Op::JMP, Ref::None, Ref::Imm(0), Ref::None, Ref::None);
                self.rec.code[jz].b = Ref::Imm(self.rec.code.len() as u32);
                for x in &b.else_branch { self.stmt(x)?; }
                self.rec.code[jmp].a = Ref::Imm(self.rec.code.len() as u32);
            }
            Instruction::Return(r) => {                    // return_bucket.rs:70-120
                if r.with_size > 1 {                       // `return arr;`: with_size consecutive variable slots, returned in place
                    if r.with_size > 64 { return Err(()); }
                    let base = match r.value.as_ref() {
                        Instruction::Load(l) => match (&l.address_type, &l.src) {
                            (AddressType::Variable, LocationRule::Indexed { location, .. }) => match location.as_ref() {
                                Instruction::Value(v) => v.value as u32,
                                _ => return Err(()),
                            },
                            _ => return Err(()),
                        },
                        _ => return Err(()),
                    };
                    self.push(Op::RET, Ref::None, Ref::Tmp(base), Ref::Imm(r.with_size as u32), Ref::None);
                } else {
                    let v = self.value(&r.value)?;
                    self.push(Op::RET, Ref::None, v, Ref::None, Ref::None);
                }
            }
            Instruction::Assert(_) | Instruction::Log(_) => {}   // asserts inside functions abort the C++ run; here: status, see DESIGN
            _ => return Err(()),
        }
        Ok(())
    }
}
fn function_body(body: &InstructionList, producer: &CUDAProducer, file: &mut Cb2cFile, rec: &mut FunctionRecord) -> Result<(), ()> {
    let q = producer.prime_str.parse::<BigInt>().map_err(|_| {})?;
    let mut cx = FunctionCtx { producer, file, q, rec };
    for i in body { cx.stmt(i)?; }
    Ok(())
}

pub fn lower_function(f: &FunctionCodeInfo, producer: &CUDAProducer, file: &mut Cb2cFile) -> Result<FunctionRecord, ()> {
    let mut rec = FunctionRecord { name: f.header.clone(), n_params: f.params.iter().map(|p| p.length.iter().product::<usize>().max(1)).sum::<usize>() as u32,
                                   n_regs: f.max_number_of_vars as u32, code: vec![] };
    if rec.n_regs > 192 { return Err(()); }
    function_body(&f.body, producer, file, &mut rec)?;
    Ok(rec)
}

/// Circuit::produce_cuda: the counterpart of Circuit::produce_c (circuit.rs:596-612)
/// `signal_names_of(template id)`: the names of the instance's signals in numbering order, array elements spelled out
/// (from the VCP: TemplateInstance::signals, the same list `--sym` walks) - empty to omit the symbols section.
pub fn produce_cb2c(templates: &[Box<TemplateCodeInfo>], functions: &[Box<FunctionCodeInfo>], producer: &CUDAProducer,
                    constraints_of: &dyn Fn(usize) -> Vec<[Vec<(Ref, BigInt)>; 3]>,
                    signal_names_of: &dyn Fn(usize) -> Vec<String>) -> Result<Cb2cFile, ()> {
    let q = producer.prime_str.parse::<BigInt>().map_err(|_| {})?;
    let mut file = Cb2cFile::default();
    file.prime = producer.prime_id()?;
    // callee first: `producer.function_ids` numbers the functions in a topological order of the call graph (computed by
    // `Circuit::cuda_producer`; a cycle - a recursive function - has no such order and is refused there), and `functions`
    // is walked in that order so that a body only names functions that are already in `file.functions`
    let mut order: Vec<&Box<FunctionCodeInfo>> = functions.iter().collect();
    order.sort_by_key(|f| producer.function_ids.get(&f.header).copied().unwrap_or(u32::MAX));
    for f in order { let r = lower_function(f, producer, &mut file)?; file.functions.push(r); }
    for t in templates {
        if t.is_extern_c { return Err(()); }
        let mut cx = TemplateCtx { producer, file: &mut file, q: q.clone(), rec: TemplateRecord::default(),
                                   vars: vec![Val::Known(BigInt::from(0)); t.var_stack_depth], sub_of_cmp: vec![],
                                   n_out: t.number_of_outputs, n_in: t.number_of_inputs };
        cx.rec.name = t.header.clone();
        cx.rec.n_out = t.number_of_outputs as u32;
        cx.rec.n_in = t.number_of_inputs as u32;
        cx.rec.n_inter = t.number_of_intermediates as u32;
        for i in &t.body { i.produce_cuda(&mut cx)?; }
        let mut rec = cx.rec;
        for con in constraints_of(t.id) {                            // ConstraintExporter view of this template instance
            let mut row: [Vec<(Ref, u32)>; 3] = Default::default();
            for (k, lc) in con.iter().enumerate() { for (r, c) in lc { row[k].push((*r, file.const_id(c, &q))); } }
            rec.constraints.push(row);
        }
        rec.signal_names = signal_names_of(t.id);
        file.templates.push(rec);
    }
    file.main = (templates.len() - 1) as u32;                        // the main template is instantiated last
    for info in &producer.main_input_list { file.names.push((info.name.clone(), info.start as u32, info.size as u32)); }
    Ok(file)
}
