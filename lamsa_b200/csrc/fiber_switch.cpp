// fiber_switch.cpp -- the context switch of the batch producer's worker fibers (fiber_sched.cu).
// glibc's swapcontext saves and restores the signal mask with a system call on every switch; a read issues
// tens of DP calls and every call is two switches, so the producer makes several hundred thousand switches
// per chunk of reads.  This switch saves what the System V x86-64 ABI requires of a callee (rbx, rbp,
// r12-r15, the stack pointer, MXCSR and the x87 control word) and nothing else.  Plain host C++ with a
// top-level asm block; other architectures use ucontext (LB2_FIBER_UCONTEXT in fiber_sched.cu).
#if defined(__x86_64__)
extern "C" void lb2_fiber_swap(void** save_sp, void* load_sp);
extern "C" void lb2_fiber_entry_thunk(void);
extern "C" void lb2_fiber_main(void* fiber);      // fiber_sched.cu; never returns

asm(R"(
    .text
    .globl lb2_fiber_swap
    .type lb2_fiber_swap, @function
lb2_fiber_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size lb2_fiber_swap, .-lb2_fiber_swap

    .globl lb2_fiber_entry_thunk
    .type lb2_fiber_entry_thunk, @function
lb2_fiber_entry_thunk:
    movq %r12, %rdi
    call lb2_fiber_main@PLT
    ud2
    .size lb2_fiber_entry_thunk, .-lb2_fiber_entry_thunk
    .text
)");
#endif
